"""A CPU stand-in for snag_b200.ops built on the oracle, with the same five entry points the sharded
evaluation driver uses. TEST-ONLY: it lets the partition / candidate-merge / counter-reduction logic of
snag_b200.evaluate run under gloo (or the in-process lockstep simulator) without a GPU."""
from __future__ import annotations

import numpy as np
import torch

from oracle import oracle

KT = 16


def _np(t):
    return t.detach().cpu().float().numpy()


def _c_matrix(X, Y, xn, yn, n1, n2):
    s = oracle.dot_matrix(_np(X[:n1]), _np(Y[:n2]))
    t = (_np(xn)[:n1, None] + _np(yn)[None, :n2]).astype(np.float32)
    d = np.maximum(t - np.float32(2.0) * s, np.float32(0.0)).astype(np.float32)
    return d


def eval_rowtopk(X, Y, xn, yn, n1, n2, want_idx=False):
    c = (np.float32(1.0) - _c_matrix(X, Y, xn, yn, n1, n2)).astype(np.float32)
    part = np.full((1, n1, KT), -np.inf, np.float32)
    pidx = np.full((1, n1, KT), -1, np.int32)
    kk = min(KT, n2)
    order = np.lexsort((-np.broadcast_to(np.arange(n2), c.shape), c), axis=1)[:, -kk:]   # ascending, lower id wins ties
    part[0, :, KT - kk:] = np.take_along_axis(c, order, 1)
    pidx[0, :, KT - kk:] = order
    if want_idx:
        return torch.from_numpy(part), torch.from_numpy(pidx)
    return torch.from_numpy(part)


def topk_merge_mean(part, k, want_nv=True, want_cand=False, part_idx=None):
    p = part.numpy()
    allv = np.concatenate(list(p), axis=1)
    if part_idx is not None:
        alli = np.concatenate(list(part_idx.numpy()), axis=1)
        order = np.lexsort((-alli, allv), axis=1)[:, -KT:]
        cand_i = np.take_along_axis(alli, order, 1)
        allv = np.take_along_axis(allv, order, 1)
    else:
        allv = np.sort(allv, axis=1)[:, -KT:]        # ascending, KT largest
    nv = None
    if want_nv:
        s = np.zeros((allv.shape[0],), np.float32)
        for t in range(k):
            s = (s + allv[:, KT - 1 - t]).astype(np.float32)
        nv = torch.from_numpy((s / np.float32(k)).astype(np.float32))
    if part_idx is not None:
        return nv, torch.from_numpy(allv.copy()), torch.from_numpy(cand_i.astype(np.int32))
    cand = torch.from_numpy(allv.copy()) if want_cand else None
    return nv, cand


def topk_rescore(A, B, an, bn, cand_idx, cand_val, k, n_b, tag="rows", want_best=False, outsider_bound=None):
    """The stand-in's scores are already canonical: the mean of the k largest candidate values, largest first; the
    nearest candidate is the one with the smallest canonical distance, lowest id on ties."""
    v = np.sort(cand_val.numpy(), axis=1)
    s = np.zeros((v.shape[0],), np.float32)
    for t in range(k):
        s = (s + v[:, KT - 1 - t]).astype(np.float32)
    nv = torch.from_numpy((s / np.float32(k)).astype(np.float32))
    if not want_best:
        return nv
    n_a = cand_idx.shape[0]
    d = _c_matrix(A, B, an, bn, n_a, n_b)
    ci = cand_idx.numpy().astype(np.int64)
    ok = ci >= 0
    dv = np.where(ok, np.take_along_axis(d, np.where(ok, ci, 0), 1), np.inf).astype(np.float32)
    order = np.lexsort((np.where(ok, ci, 1 << 40), dv), axis=1)[:, 0]
    rows = np.arange(n_a)
    return nv, torch.from_numpy(dv[rows, order]), torch.from_numpy(ci[rows, order].astype(np.int32))


def _dist(X, Y, xn, yn, nv1, nv2, n1, n2, use_csls):
    d = _c_matrix(X, Y, xn, yn, n1, n2)
    if not use_csls:
        return d
    c = (np.float32(1.0) - d).astype(np.float32)
    u = (np.float32(2.0) * c - _np(nv1)[:n1, None]).astype(np.float32)
    v = (u - _np(nv2)[None, :n2]).astype(np.float32)
    return (np.float32(1.0) - v).astype(np.float32)


def pair_score(X, Y, n, xn, yn, nv1, nv2, use_csls, want_dot=False):
    x, y = _np(X[:n]), _np(Y[:n])
    s = np.einsum("ij,ij->i", x.astype(np.float64), y.astype(np.float64)).astype(np.float32)
    t = (_np(xn)[:n] + _np(yn)[:n]).astype(np.float32)
    d = np.maximum(t - np.float32(2.0) * s, np.float32(0.0)).astype(np.float32)
    if use_csls:
        c = (np.float32(1.0) - d).astype(np.float32)
        u = (np.float32(2.0) * c - _np(nv1)[:n]).astype(np.float32)
        d = (np.float32(1.0) - (u - _np(nv2)[:n]).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(d)


def eval_rank(X, Y, xn, yn, nv1, nv2, g_row, g_col, row_gid0, col_gid0, n1, n2, use_csls, cnt_row, cnt_col,
              want_top3=False):
    dist = _dist(X, Y, xn, yn, nv1, nv2, n1, n2, use_csls)
    gr, gc = _np(g_row)[:n1], _np(g_col)[:n2]
    ri = row_gid0 + np.arange(n1)[:, None]
    cj = col_gid0 + np.arange(n2)[None, :]
    prow = (dist < gr[:, None]) | ((dist == gr[:, None]) & (cj < ri))
    pcol = (dist < gc[None, :]) | ((dist == gc[None, :]) & (ri < cj))
    same = ri == cj
    prow &= ~same
    pcol &= ~same
    cnt_row[:n1] += torch.from_numpy(prow.sum(1).astype(np.int32))
    cnt_col[:n2] += torch.from_numpy(pcol.sum(0).astype(np.int32))
    if not want_top3:
        return None, None
    # nearest-candidate lists as the CUDA sweep emits them: 4 per row, score descending (= distance ascending)
    order = np.lexsort((np.broadcast_to(cj, dist.shape), dist), axis=1)[:, :4]
    v = np.full((1, n1, 4), -np.inf, np.float32)
    i = np.full((1, n1, 4), 0x7FFFFFFF, np.int32)
    m = order.shape[1]
    v[0, :, :m] = -np.take_along_axis(dist, order, 1)
    i[0, :, :m] = order + col_gid0
    return torch.from_numpy(v), torch.from_numpy(i)


def top4_merge(val, idx):
    v = np.concatenate(list(val.numpy()), axis=1)
    i = np.concatenate(list(idx.numpy()), axis=1)
    order = np.lexsort((i, -v), axis=1)[:, :4]
    return (torch.from_numpy(np.take_along_axis(v, order, 1).astype(np.float32)),
            torch.from_numpy(np.take_along_axis(i, order, 1).astype(np.int32)))


def top3_rescore(X, Y, xn, yn, nv1, nv2, use_csls, cand):
    n1, n2 = cand.shape[0], _np(yn).shape[0]
    dist = _dist(X, Y, xn, yn, nv1, nv2, n1, n2, use_csls)
    ci = cand.numpy().astype(np.int64)
    ok = ci != 0x7FFFFFFF
    v = np.where(ok, np.take_along_axis(dist, np.where(ok, ci, 0), 1), np.inf).astype(np.float32)
    order = np.lexsort((ci, v), axis=1)
    return (torch.from_numpy(np.take_along_axis(v, order, 1)),
            torch.from_numpy(np.take_along_axis(ci, order, 1).astype(np.int32)))


# ------------------------------------------------------------------------------------------------ ICL loss
# CPU stand-ins for the four kernel entry points the (anchor-sharded) ICL loss uses, following the formulas in
# include/snag_b200.h; fp32 torch on the bf16-rounded operands.
def round_up(x, m):
    return (x + m - 1) // m * m


def prep_bf16(emb, idx=None, normalize=True, rows_pad_to=1, out=None):
    x = emb if idx is None else emb.index_select(0, idx)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=1)
    n, d = x.shape
    if out is None:
        out = torch.zeros((round_up(n, rows_pad_to), round_up(d, 64)), dtype=torch.bfloat16)
    out[:n, :d] = x.to(torch.bfloat16)
    out[:n, d:] = 0
    return out, (out[:n].float() ** 2).sum(1)


def normalize_bwd_scatter(emb, idx, dz, demb, normalize=True):
    with torch.enable_grad():                         # called from inside an autograd backward (grad mode is off there)
        e = (emb if idx is None else emb.index_select(0, idx)).detach().clone().requires_grad_(True)
        z = torch.nn.functional.normalize(e, dim=1) if normalize else e * 1.0
        z.backward(dz[:e.shape[0]])
    if idx is None:
        demb += e.grad
    else:
        demb.index_add_(0, idx, e.grad)


def _icl_logits(X, Y, B, Bp, inv_tau, row0, nx):
    s = X[:nx].float() @ Y.float().t()                                   # [nx, 2Bp]
    col = torch.arange(2 * Bp)
    part, idx = col // Bp, col % Bp
    gr = row0 + torch.arange(nx)
    dead = (idx >= B)[None, :] | ((part == 1)[None, :] & (idx[None, :] == gr[:, None]))
    return s, dead, gr


def icl_side(X, Y, B, Bp, inv_tau, row0=0, nx=None):
    nx = Bp if nx is None else nx
    s, dead, gr = _icl_logits(X, Y, B, Bp, inv_tau, row0, nx)
    valid = max(0, min(nx, B - row0))
    logits = (s * inv_tau).masked_fill(dead, float("-inf"))[:valid]
    lse = torch.logsumexp(logits, 1)
    pos = s[torch.arange(valid), gr[:valid]]
    return lse, lse - pos * inv_tau, pos


ICL_SYM_MAX_PROBLEMS = 16


def icl_fwd_sym(S3s, B, Bp, inv_tau, rank=0, world=1, all_reduce=None, esave=None):
    """CPU stand-in for ops.icl_fwd_sym: the strict upper triangle of Z.Z^T (Z = [a ; b]) contributes every element to the
    row sum of its row and of its column; with world > 1 this rank takes the elements whose ROW index is congruent to its
    rank (a different split of the same work than the kernel's unit ranges — any split must add up) and `all_reduce` sums
    the partial totals and the positive logits. Returns [n_prob, 4, B] = (lse_a, nll_a, lse_b, nll_b)."""
    outs = []
    parts = []
    for S3 in S3s:
        Z = S3[:2 * Bp].float()
        S = Z @ Z.t()
        valid = torch.zeros(2 * Bp, dtype=torch.bool)
        valid[:B] = True
        valid[Bp:Bp + B] = True
        E = torch.exp(S * inv_tau - inv_tau)
        keep = torch.triu(torch.ones_like(S, dtype=torch.bool), 1) & valid[:, None] & valid[None, :]
        mine = (torch.arange(2 * Bp) % world == rank)[:, None]
        E = torch.where(keep & mine, E, torch.zeros(()))
        total = E.sum(1) + E.sum(0)
        pos = torch.zeros(Bp)
        own = (torch.arange(B) % world == rank)
        idx = torch.arange(B)[own]
        pos[idx] = S[idx, Bp + idx]
        parts.append(torch.cat([total, pos]))
    buf = torch.stack(parts, 0).reshape(-1).contiguous()
    if world > 1:
        buf = all_reduce(buf)
    buf = buf.view(len(S3s), 3 * Bp)
    for p in range(len(S3s)):
        total, pos = buf[p, :2 * Bp], buf[p, 2 * Bp:]
        la = torch.log(total[:B]) + inv_tau
        lb = torch.log(total[Bp:Bp + B]) + inv_tau
        ps = pos[:B] * inv_tau
        outs.append(torch.stack([la, la - ps, lb, lb - ps], 0))
    return torch.stack(outs, 0)


def icl_bwd_logits(X, Y, B, Bp, inv_tau, cr, cc, dg, row0=0, nx=None, self_cols=True, ebar=0.0):
    nx = Bp if nx is None else nx
    s, dead, gr = _icl_logits(X, Y, B, Bp, inv_tau, row0, nx)
    E = torch.exp(s * inv_tau - inv_tau) - ebar
    ok = gr < B
    crp = torch.zeros(Bp); crp[:B] = cr[:B]
    ccp = torch.zeros(Bp); ccp[:B] = cc[:B]
    dgp = torch.zeros(Bp); dgp[:B] = dg[:B]
    cr_i = torch.where(ok, crp[gr.clamp(max=Bp - 1)], torch.zeros(()))
    colc = torch.cat([ccp, crp if self_cols else torch.zeros(Bp)])       # part 0: cc_j, part 1: cr_j (or nothing)
    G = (cr_i[:, None] + colc[None, :]) * E * inv_tau
    i = torch.arange(nx)[ok]
    G[i, gr[ok]] -= dgp[gr[ok]] * inv_tau                                # positive logit of part 0
    G = G.masked_fill(dead, 0.0)
    G[~ok] = 0.0
    return G.to(torch.bfloat16)


def contract(P, Q, n1, n2):
    return P[:n1].float() @ Q[:n2].float().t()


# ------------------------------------------------------------------------------------------------ link mining
def grad_contract(G, YT, n_rows, d):
    return contract(G, YT, n_rows, d)


def mutual_nn(X, Y, xn, yn, n1, n2, colb):
    """CPU stand-in for ops.mutual_nn: one list; columns only where the pre-filter admits their minimum (as the kernel)."""
    d = _c_matrix(X, Y, xn, yn, n1, n2)
    s = oracle.dot_matrix(_np(X[:n1]), _np(Y[:n2]))
    row_val = torch.from_numpy(d.min(1)[None, :].copy())
    row_idx = torch.from_numpy(d.argmin(1).astype(np.int32)[None, :].copy())
    flagged = s > (np.float32(0.5) * _np(xn)[:n1, None] + _np(colb)[None, :n2])
    key = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(n1, dtype=np.uint64)[:, None]
    key = np.where(flagged, key, np.uint64(0xFFFFFFFFFFFFFFFF)).min(0)
    return row_val, row_idx, torch.from_numpy(key.view(np.int64).copy())
