"""Test / benchmark harness around the UNMODIFIED reference (zjukg/SNAG, SNAG_MMEA): where it lives on this machine,
how to import it, a synthetic dataset in the reference's own on-disk format, and its command line.

Test infrastructure only — nothing under snag_b200/ imports this module. The reference is imported in place from
baseline/_ref/SNAG_MMEA (the copy baseline/install_ref.py makes; it travels to the GPU box) or, in the build
container, from /root/reference/SNAG_MMEA. Nothing of it is copied into the repository.

Synthetic data: two knowledge graphs of `n_side` entities each, `n_links` aligned pairs (entity i of
KG 1 <-> entity n_side + perm[i] of KG 2), ~5 random triples per entity over 200 relations, Bernoulli attribute sets,
image vectors for 6/7 of the entities. Aligned entities get correlated features (shared neighbours through the
links, overlapping attribute sets, noisy copies of the same image vector) so that a short training run actually
learns an alignment and the evaluation has non-trivial ranks.
"""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(HERE, "stubs")
_CANDIDATES = (os.path.join(HERE, "_ref", "SNAG_MMEA"), "/root/reference/SNAG_MMEA")


def ref_root() -> str | None:
    """Directory of the reference's SNAG_MMEA tree on this machine, or None."""
    for p in _CANDIDATES:
        if os.path.isfile(os.path.join(p, "main.py")):
            return p
    return None


def load_reference():
    """Make `import model`, `import src`, `import main` resolve to the unmodified reference. On a machine without a
    GPU the hard-coded `.cuda()` calls of the reference (model/SNAG_loss.py:90,96,165; model/SNAG.py:23-36) become
    the identity so that its CPU path can be imported and run for goldens. Returns the reference root."""
    import torch
    root = ref_root()
    if root is None:
        raise RuntimeError("no reference checkout: run baseline/install_ref.py where /root/reference exists")
    sys.dont_write_bytecode = True
    for p in (STUBS, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    return root


# ================================================================================================ synthetic dataset
def synth_graph(n_side: int, n_links: int, seed: int = 3408, img_dim: int = 256, n_rel: int = 200, n_attr: int = 300,
                triples_per_ent: int = 5):
    """The raw synthetic graph pair as Python objects (ids as in the reference's files: KG 1 = 0..n_side-1,
    KG 2 = n_side..2 n_side-1)."""
    rng = np.random.RandomState(seed)
    n_ent = 2 * n_side
    left = rng.permutation(n_side)[:n_links]
    right = n_side + rng.permutation(n_side)[:n_links]
    partner = {int(a): int(b) for a, b in zip(left, right)}
    # KG 1 triples at random; KG 2 mirrors those whose endpoints are both aligned (shared structure), plus its own
    t1 = [(int(h), int(r), int(t)) for h, r, t in zip(rng.randint(0, n_side, triples_per_ent * n_side),
                                                       rng.randint(0, n_rel, triples_per_ent * n_side),
                                                       rng.randint(0, n_side, triples_per_ent * n_side))]
    t2 = [(partner[h], r + n_rel, partner[t]) for h, r, t in t1 if h in partner and t in partner and rng.rand() < 0.8]
    extra = triples_per_ent * n_side - len(t2)
    t2 += [(int(h), int(r), int(t)) for h, r, t in zip(n_side + rng.randint(0, n_side, extra),
                                                       n_rel + rng.randint(0, n_rel, extra),
                                                       n_side + rng.randint(0, n_side, extra))]
    # attributes: every entity draws ~3 of n_attr; an aligned partner copies most of them
    attrs = {}
    for e in range(n_side):
        attrs[e] = set(rng.choice(n_attr, size=1 + rng.randint(0, 5), replace=False).tolist())
    for e in range(n_side, n_ent):
        attrs[e] = set(rng.choice(n_attr, size=1 + rng.randint(0, 5), replace=False).tolist())
    for a, b in partner.items():
        attrs[b] = set(x for x in attrs[a] if rng.rand() < 0.8) | set(x for x in attrs[b] if rng.rand() < 0.3)
        if not attrs[b]:
            attrs[b] = {int(rng.randint(0, n_attr))}
    # images: 6/7 of the entities have one; partners see a noisy copy
    has_img = rng.rand(n_ent) < 6.0 / 7.0
    img = rng.randn(n_ent, img_dim).astype(np.float32)
    for a, b in partner.items():
        img[b] = img[a] + 0.5 * rng.randn(img_dim).astype(np.float32)
    img_dict = {int(e): img[e] for e in range(n_ent) if has_img[e]}
    ills = [(int(a), int(b)) for a, b in zip(left, right)]
    return dict(n_side=n_side, n_ent=n_ent, ills=ills, triples_1=t1, triples_2=t2, attrs=attrs, img_dict=img_dict)


def write_dataset(data_path: str, n_side: int = 600, n_links: int = 400, seed: int = 3408, img_dim: int = 256,
                  split: str = "ja_en") -> dict:
    """Write the synthetic graph pair under `data_path` in the layout load_eva_data (src/data.py:135-272) reads for
    --data_choice DBP15K --data_split `split`: ent_ids_{1,2}, ill_ent_ids, triples_{1,2}, training_attrs_{1,2} and
    pkls/<split>_GA_id_img_feature_dict.pkl. Returns the graph dict."""
    g = synth_graph(n_side, n_links, seed, img_dim)
    d = os.path.join(data_path, "DBP15K", split)
    os.makedirs(d, exist_ok=True)
    os.makedirs(os.path.join(data_path, "pkls"), exist_ok=True)
    name = lambda e: f"http://synthetic/resource/E{e}"
    for side, rng_ in ((1, range(0, n_side)), (2, range(n_side, 2 * n_side))):
        with open(os.path.join(d, f"ent_ids_{side}"), "w", encoding="utf-8") as f:
            for e in rng_:
                f.write(f"{e}\t{name(e)}\n")
        with open(os.path.join(d, f"training_attrs_{side}"), "w", encoding="utf-8") as f:
            for e in rng_:
                f.write("\t".join([name(e)] + [f"attr{a}" for a in sorted(g["attrs"][e])]) + "\n")
        with open(os.path.join(d, f"triples_{side}"), "w", encoding="utf-8") as f:
            for h, r, t in g[f"triples_{side}"]:
                f.write(f"{h}\t{r}\t{t}\n")
    with open(os.path.join(d, "ill_ent_ids"), "w", encoding="utf-8") as f:
        for a, b in g["ills"]:
            f.write(f"{a}\t{b}\n")
    with open(os.path.join(data_path, "pkls", f"{split}_GA_id_img_feature_dict.pkl"), "wb") as f:
        pickle.dump(g["img_dict"], f)
    return g


def main_argv(data_path: str, epochs: int = 2, batch_size: int = 128, extra: list[str] | None = None) -> list[str]:
    """The scripted SNAG command line (run_snag.sh:2-45) scaled down to a synthetic dataset: same switches, small
    widths, `epochs` epochs with an evaluation after each, final test + prediction file."""
    argv = ["--gpu", "0", "--eval_epoch", "1", "--only_test", "0", "--model_name", "SNAG", "--data_choice", "DBP15K",
            "--data_split", "ja_en", "--data_rate", "0.3", "--epoch", str(epochs), "--lr", "5e-4",
            "--hidden_units", "64,64,64", "--save_model", "0", "--batch_size", str(batch_size), "--semi_learn_step", "5",
            "--csls", "--csls_k", "3", "--random_seed", "3408", "--exp_name", "snag_b200_e2e", "--exp_id", "e2e",
            "--workers", "1", "--accumulation_steps", "1", "--scheduler", "cos", "--attr_dim", "64", "--img_dim", "64",
            "--name_dim", "64", "--char_dim", "64", "--hidden_size", "64", "--intermediate_size", "128", "--tau", "0.1",
            "--tau2", "4.0", "--structure_encoder", "gat", "--num_attention_heads", "1", "--num_hidden_layers", "1",
            "--use_surface", "0", "--use_intermediate", "1", "--replay", "0", "--ratio", "1.0", "--add_noise", "1",
            "--noise_ratio", "0.2", "--mask_ratio", "0.7", "--data_path", os.path.abspath(data_path)]
    # (no --no_tensorboard: Runner.train calls self.writer.add_scalars unconditionally, main.py:283)
    return argv + list(extra or [])


def parse_args(argv: list[str]):
    """The reference's own argument parser and post-processing (config.py:8-218) applied to `argv`."""
    load_reference()
    import config as ref_config
    saved = sys.argv
    sys.argv = ["main.py"] + list(argv)
    try:
        c = ref_config.cfg()
        c.get_args()
        return c.update_train_configs()
    finally:
        sys.argv = saved
