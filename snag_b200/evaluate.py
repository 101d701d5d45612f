"""Alignment evaluation: source x target similarity, CSLS correction, ranking into Hits@k / MR / MRR.

Host-side mirror of the reference's evaluation interface
  - pairwise_distances(x, y=None)          SNAG_MMEA/src/utils.py:202-218
  - csls_sim(sim_mat, k)                   SNAG_MMEA/src/utils.py:417-435
  - the ranking body of Runner._test       SNAG_MMEA/main.py:379-436
on top of the fused tcgen05 sweeps of libsnag_b200.so. The n x n matrix is never materialised on the
fused path (`align_ranks`); the two drop-ins that must return a matrix do materialise it.

Sharded evaluation (targets split across ranks, one process per GPU) lives in `align_ranks` too: pass a
torch.distributed process group. Per-row candidates / counters are exchanged with NCCL all-gather /
all-reduce; integer counters make the result identical for any number of ranks.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import ops
from ._lib import KT, SnagError

TOP_K = (1, 10, 50)   # main.py:380


# =================================================================================================
# fused path
# =================================================================================================
@dataclass
class AlignRanks:
    rank_l2r: torch.Tensor            # int32 [n]  0-based position of target i in the sorted row i
    rank_r2l: torch.Tensor            # int32 [n]  0-based position of source j in the sorted column j
    nv1: torch.Tensor | None          # fp32 [n]   CSLS row neighbourhood means
    nv2: torch.Tensor | None          # fp32 [n]   CSLS column neighbourhood means
    g: torch.Tensor                   # fp32 [n]   distance of the ground-truth pair
    top3_idx: torch.Tensor | None = None   # int32 [n,3] pair ids of the 3 nearest targets per source
    top3_val: torch.Tensor | None = None
    launches: int = 0                 # kernels of libsnag_b200.so launched
    info: dict = field(default_factory=dict)


def shard_bounds(n: int, world: int, rank: int, align: int = 256) -> tuple[int, int]:
    """Contiguous target shard [c0, c1) of rank `rank`; shard size is a multiple of `align` except the last."""
    per = ops.round_up((n + world - 1) // world, align)
    c0 = min(rank * per, n)
    c1 = min(c0 + per, n)
    return c0, c1


def align_ranks(X: torch.Tensor, Y: torch.Tensor, xn: torch.Tensor, yn: torch.Tensor, n: int, csls_k: int = 10,
                use_csls: bool = True, want_top3: bool = False, group=None) -> AlignRanks:
    """Fused evaluation of n aligned pairs (x_i <-> y_i).

    X, Y : bf16 operands [>=n, Dpad] from ops.prep_bf16; xn, yn : their squared norms [n].
    With `group` (torch.distributed), every rank holds all of X and Y and sweeps only its target shard.
    """
    if use_csls and not 1 <= csls_k <= KT:
        raise SnagError(f"csls_k={csls_k} unsupported: the fused CSLS path keeps {KT} candidates per row")
    if use_csls and csls_k > n:
        # torch.topk raises in the reference (src/utils.py:431) when k exceeds the matrix side
        raise ValueError(f"csls_k={csls_k} exceeds the number of evaluated pairs n={n}")
    dev = X.device
    launches = 0
    if group is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        dist = None
        world, rank = 1, 0
    c0, c1 = shard_bounds(n, world, rank)
    ns = c1 - c0                                   # targets owned by this rank
    per = shard_bounds(n, world, 0)[1]             # padded shard size used for the gathers
    Ys = Y[c0:c1] if ns > 0 else None
    yns = yn[c0:c1] if ns > 0 else None

    nv1 = nv2 = None
    if use_csls:
        # sweep 1: row neighbourhoods (every source against this rank's targets)
        if ns > 0:
            part = ops.eval_rowtopk(X, Ys, xn, yns, n, ns)
            launches += 1
        else:
            part = torch.full((1, n, KT), float("-inf"), dtype=torch.float32, device=dev)
        if world == 1:
            nv1, _ = ops.topk_merge_mean(part, csls_k)
            launches += 1
        else:
            _, cand = ops.topk_merge_mean(part, csls_k, want_nv=False, want_cand=True)
            allc = torch.empty((world, n, KT), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allc, cand, group=group)
            nv1, _ = ops.topk_merge_mean(allc, csls_k)
            launches += 2
        # sweep 1': column neighbourhoods (this rank's targets against every source)
        nv2_loc = torch.zeros((per,), dtype=torch.float32, device=dev)
        if ns > 0:
            part2 = ops.eval_rowtopk(Ys, X, yns, xn, ns, n)
            nv2s, _ = ops.topk_merge_mean(part2, csls_k)
            nv2_loc[:ns] = nv2s
            launches += 2
        if world == 1:
            nv2 = nv2_loc[:n]
        else:
            allv = torch.empty((world * per,), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allv, nv2_loc, group=group)
            nv2 = allv[:n].contiguous()

    # ground-truth scores of all pairs (n dot products; replicated on every rank)
    g = ops.pair_score(X, Y, n, xn, yn, nv1, nv2, use_csls)
    launches += 1

    # sweep 2: rank counters
    cnt_row = torch.zeros((n,), dtype=torch.int32, device=dev)
    cnt_col_loc = torch.zeros((per,), dtype=torch.int32, device=dev)
    t3v = t3i = None
    if ns > 0:
        nv2s = nv2[c0:c1] if use_csls else None
        t3v, t3i = ops.eval_rank(X, Ys, xn, yns, nv1, nv2s, g, g[c0:c1], 0, c0, n, ns, use_csls, cnt_row, cnt_col_loc,
                                 want_top3)
        launches += 1
    top3_idx = top3_val = None
    if want_top3:
        if ns > 0:
            t3v, t3i = ops.top3_merge(t3v, t3i)
            launches += 1
        else:
            t3v = torch.full((n, 4), float("inf"), dtype=torch.float32, device=dev)
            t3i = torch.full((n, 4), 0x7FFFFFFF, dtype=torch.int32, device=dev)
    if world == 1:
        rank_l2r, rank_r2l = cnt_row, cnt_col_loc[:n]
        if want_top3:
            top3_idx, top3_val = t3i[:, :3], t3v[:, :3]
    else:
        dist.all_reduce(cnt_row, group=group)
        allc = torch.empty((world * per,), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allc, cnt_col_loc, group=group)
        rank_l2r, rank_r2l = cnt_row, allc[:n].contiguous()
        if want_top3:
            gv = torch.empty((world, n, 4), dtype=torch.float32, device=dev)
            gi = torch.empty((world, n, 4), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gv, t3v.contiguous(), group=group)
            dist.all_gather_into_tensor(gi, t3i.contiguous(), group=group)
            t3v, t3i = ops.top3_merge(gv, gi)
            launches += 1
            top3_idx, top3_val = t3i[:, :3], t3v[:, :3]
    return AlignRanks(rank_l2r, rank_r2l, nv1, nv2, g, top3_idx, top3_val, launches,
                      {"world": world, "shard": (c0, c1)})


# =================================================================================================
# metrics (host glue; follows main.py:380-436 operation by operation)
# =================================================================================================
@dataclass
class AlignMetrics:
    acc: np.ndarray      # np.float32 [3]  Hits@1/10/50, rounded to 4 decimals like the reference
    mr: float            # mean rank, 1-based
    mrr: float
    hits: np.ndarray     # int64 [3] raw counts


def metrics_from_ranks(ranks, top_k=TOP_K) -> AlignMetrics:
    """Hits@k / MR / MRR from 0-based integer ranks, reproducing the reference's accumulators:
    hit counters are np.float32 (main.py:382-383), mean uses Python ints, mrr is a Python float summed in
    index order (main.py:403-404) -> np.cumsum in float64 is the same sequence of additions."""
    r = np.asarray(ranks.detach().cpu().numpy() if isinstance(ranks, torch.Tensor) else ranks).astype(np.int64)
    n = r.shape[0]
    if n == 0:
        raise ValueError("no test pairs")
    hits = np.array([(r < k).sum() for k in top_k], dtype=np.int64)
    acc = np.zeros((len(top_k),), dtype=np.float32)
    for i in range(len(top_k)):
        acc[i] = round(np.float32(hits[i]) / n, 4)
    mr = float(int((r + 1).sum())) / n
    mrr = float(np.cumsum(1.0 / (r + 1).astype(np.float64))[-1]) / n
    return AlignMetrics(acc, mr, mrr, hits)


def evaluate_alignment(final_emb: torch.Tensor, test_left: torch.Tensor, test_right: torch.Tensor, csls: bool = True,
                       csls_k: int = 10, want_top3: bool = False, normalize: bool = True, group=None) -> dict:
    """The evaluation a user of the reference gets from Runner._test, as a function:
    final_emb [N, D] fp32 (device), test_left/right LongTensor [n] -> metrics for both directions."""
    n = test_left.numel()
    if test_right.numel() != n:
        raise ValueError("test_left and test_right must pair up")
    X, xn = ops.prep_bf16(final_emb, test_left.to(torch.int64).contiguous(), normalize)
    Y, yn = ops.prep_bf16(final_emb, test_right.to(torch.int64).contiguous(), normalize)
    res = align_ranks(X, Y, xn, yn, n, csls_k, csls, want_top3, group)
    out = {
        "l2r": metrics_from_ranks(res.rank_l2r),
        "r2l": metrics_from_ranks(res.rank_r2l),
        "ranks": res,
        "launches": res.launches + 2,
    }
    return out


# =================================================================================================
# materialising drop-ins
# =================================================================================================
def pairwise_distances(x: torch.Tensor, y: torch.Tensor | None = None) -> torch.Tensor:
    """Drop-in for src/utils.py:202-218: dist[i,j] = clamp(||x_i||^2 + ||y_j||^2 - 2 x_i.y_j, 0), fp32 [n1,n2].
    Operands are rounded to bf16 for the tensor-core contraction; norms are those of the rounded rows."""
    x = x.contiguous().float()
    X, xn = ops.prep_bf16(x, None, normalize=False)
    if y is None:
        Y, yn = X, xn
        n2 = x.shape[0]
    else:
        y = y.contiguous().float()
        Y, yn = ops.prep_bf16(y, None, normalize=False)
        n2 = y.shape[0]
    return ops.sim_write(X, Y, xn, yn, x.shape[0], n2, mode=1)
