"""Dev probe: how far are the fused ial_loss and a plain fp32 torch evaluation from an fp64 evaluation of the same
bf16-rounded operands (loss and gradient)? Decides the tolerances of tests/test_loss_gpu.py."""
import json
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from snag_b200 import loss as sloss  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False


def ref(src, tar, il, ir, tau, alpha, zoom, red, dt):
    rnd = lambda t: F.normalize(t.float(), dim=1).to(torch.bfloat16).to(dt)
    rs, rt = rnd(src.detach()), rnd(tar)
    zs = F.normalize(src.to(dt), dim=1)
    zs = zs + (rs - zs).detach()
    s_i, s_j, t_i, t_j = zs[il], zs[ir], rt[il], rt[ir]
    B = il.numel()
    eye = torch.eye(B, device=src.device, dtype=dt) * 1e9
    cat = lambda u, v: torch.cat([u @ v.t() / tau, u @ u.t() / tau - eye], 1)
    la = F.kl_div(F.log_softmax(cat(s_i, s_j), 1), F.softmax(cat(t_i, t_j), 1), reduction="none")
    lb = F.kl_div(F.log_softmax(cat(s_j, s_i), 1), F.softmax(cat(t_j, t_i), 1), reduction="none")
    la, lb = (la.mean(), lb.mean()) if red == "mean" else (la.sum(), lb.sum())
    return zoom * (alpha * la + (1 - alpha) * lb)


rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for B, Ds, Dt, tau, red in [(700, 96, 320, 0.5, "mean"), (1000, 300, 1200, 4.0, "sum"), (130, 64, 64, 0.2, "mean"),
                            (3500, 300, 1200, 4.0, "mean")]:
    g = torch.Generator(device="cuda").manual_seed(B + Ds)
    N = 2 * B + 50
    base = torch.randn((N, Dt), generator=g, device=dev)
    tar = base + 0.3 * torch.randn((N, Dt), generator=g, device=dev)
    src0 = base[:, :Ds] + 0.5 * torch.randn((N, Ds), generator=g, device=dev)
    perm = torch.randperm(N, generator=g, device=dev)
    links = torch.stack([perm[:B], perm[B:2 * B]], 1)
    il, ir = sloss._links_to_index(links, dev)
    crit = sloss.ial_loss(tau=tau, ab_weight=0.3, zoom=0.1, reduction=red)
    a = src0.clone().requires_grad_(True)
    fused = crit(a, tar, links)
    fused.backward()
    outs = {}
    for name, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
        b = src0.clone().requires_grad_(True)
        r = ref(b, tar, il, ir, tau, 0.3, 0.1, red, dt)
        r.backward()
        outs[name] = (float(r), b.grad.clone())
    l64, g64 = outs["fp64"]
    print(json.dumps({"B": B, "Ds": Ds, "Dt": Dt, "tau": tau, "red": red, "loss64": l64,
                      "fused_loss_rel": abs(float(fused) - l64) / abs(l64), "fp32_loss_rel": abs(outs["fp32"][0] - l64) / abs(l64),
                      "fused_grad_rel": rel(a.grad, g64), "fp32_grad_rel": rel(outs["fp32"][1], g64)}), flush=True)

# goldens: the reference's own fp32 outputs on fp32 inputs (the fused path rounds the normalised rows to bf16)
import os  # noqa: E402
GD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
for f in sorted(os.listdir(GD)):
    if not (f.startswith("ial_") and f.endswith(".npz")):
        continue
    with np.load(os.path.join(GD, f)) as z:
        fx = {k: z[k] for k in z.files}
    src = torch.from_numpy(fx["src"]).to(dev).requires_grad_(True)
    tar = torch.from_numpy(fx["tar"]).to(dev)
    crit = sloss.ial_loss(tau=float(fx["tau"]), ab_weight=float(fx["ab_weight"]), zoom=float(fx["zoom"]), reduction=str(fx["reduction"]))
    out = crit(src, tar, fx["links"])
    out.backward()
    # the same reference arithmetic in fp64 on the fp32 inputs and on bf16-rounded normalised rows
    il, ir = sloss._links_to_index(fx["links"], dev)
    outs = {}
    for name, rounded in (("fp64_fp32rows", False), ("fp64_bf16rows", True)):
        b = torch.from_numpy(fx["src"]).to(dev).requires_grad_(True)
        if rounded:
            r = ref(b, tar, il, ir, float(fx["tau"]), float(fx["ab_weight"]), float(fx["zoom"]), str(fx["reduction"]), torch.float64)
        else:
            zs = F.normalize(b.double(), dim=1)
            rt = F.normalize(tar.double(), dim=1)
            tau = float(fx["tau"])
            B = il.numel()
            eye = torch.eye(B, device=dev, dtype=torch.float64) * 1e9
            cat = lambda u, v: torch.cat([u @ v.t() / tau, u @ u.t() / tau - eye], 1)
            la = F.kl_div(F.log_softmax(cat(zs[il], zs[ir]), 1), F.softmax(cat(rt[il], rt[ir]), 1), reduction="none")
            lb = F.kl_div(F.log_softmax(cat(zs[ir], zs[il]), 1), F.softmax(cat(rt[ir], rt[il]), 1), reduction="none")
            red = str(fx["reduction"])
            la, lb = (la.mean(), lb.mean()) if red == "mean" else (la.sum(), lb.sum())
            r = float(fx["zoom"]) * (float(fx["ab_weight"]) * la + (1 - float(fx["ab_weight"])) * lb)
        r.backward()
        outs[name] = (float(r), b.grad.clone())
    print(json.dumps({"golden": f, "shape": list(fx["src"].shape), "golden_loss": float(fx["loss"]),
                      "fused_vs_golden_loss": abs(float(out) - float(fx["loss"])) / abs(float(fx["loss"])),
                      "fused_vs_golden_grad": rel(src.grad.cpu(), torch.from_numpy(fx["grad_src"])),
                      "bf16rows_fp64_vs_golden_loss": abs(outs["fp64_bf16rows"][0] - float(fx["loss"])) / abs(float(fx["loss"])),
                      "bf16rows_fp64_vs_golden_grad": rel(outs["fp64_bf16rows"][1].cpu(), torch.from_numpy(fx["grad_src"])),
                      "fp32rows_fp64_vs_golden_loss": abs(outs["fp64_fp32rows"][0] - float(fx["loss"])) / abs(float(fx["loss"]))}), flush=True)
