"""Typed Python wrappers over the C ABI (one function per entry point of include/gdmcf_sm100.h).

Tensors are only containers for device memory; all arithmetic happens in libgdmcf_sm100.so.
Padded matrices: bf16 operands live in `[rows, ld]` tensors with `ld = round_up(cols, 64)` so every
TMA row stride is 16-byte aligned and K tails are zero.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_TANH, EPI_BIAS_ACT, EPI_COSINE, EPI_STORE, Epilogue, GemmDesc,
                   check, load, ptr, require_cuda, stream)

__all__ = [
    "ACT_NONE", "ACT_RELU", "ACT_TANH", "EPI_BIAS_ACT", "EPI_COSINE", "EPI_STORE", "round_up", "Bf16Mat",
    "SpmmPlan", "spmm_plan", "spmm_csr", "lightgcn_propagate", "build_norm_adj", "gemm", "cast_bf16",
    "cast_bf16_transpose", "densify_rows", "qsample_dropout", "onehot_noise", "onehot_tables",
    "encode_onehot_gather", "time_bias_table", "bias_act_rows", "gather_rows", "sgemm_small", "colsum_f32", "mix_rownorm", "row_inv_norm", "mask_topk", "topn_metrics", "colsum_f64",
    "mse_rows", "adamw_fused",
]


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class Bf16Mat:
    """bf16 operand [rows, cols] stored with leading dimension ld; `lo` holds the residual for fp32 mode."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor]
    rows: int
    cols: int

    @property
    def ld(self) -> int:
        return self.hi.shape[1]

    @staticmethod
    def empty(rows: int, cols: int, device, with_lo: bool = False, zero: bool = True) -> "Bf16Mat":
        ld = round_up(cols, 64)
        mk = torch.zeros if zero else torch.empty
        hi = mk(rows, ld, dtype=torch.bfloat16, device=device)
        lo = mk(rows, ld, dtype=torch.bfloat16, device=device) if with_lo else None
        return Bf16Mat(hi, lo, rows, cols)

    def float(self) -> torch.Tensor:
        """Debug/test helper: reconstruct fp32 [rows, cols]."""
        x = self.hi[:, : self.cols].float()
        if self.lo is not None:
            x = x + self.lo[:, : self.cols].float()
        return x


# ----------------------------------------------------------------------------------------------
# SpMM
# ----------------------------------------------------------------------------------------------
@dataclass
class SpmmPlan:
    items: torch.Tensor       # int32 [n_items, 4]
    long_rows: torch.Tensor   # int32 [n_long, 3]
    n_items: int
    n_long: int
    n_slots: int
    n_rows: int
    chunk: int


def spmm_plan(rowptr, chunk: int = 256, device="cuda") -> SpmmPlan:
    """Host-side work split (gdmcf_spmm_plan). rowptr: int32 numpy/CPU tensor [n_rows + 1]."""
    lib = load()
    rp = np.ascontiguousarray(rowptr.cpu().numpy() if isinstance(rowptr, torch.Tensor) else rowptr, dtype=np.int32)
    n_rows = rp.shape[0] - 1
    ni, nl, ns = C.c_int(0), C.c_int(0), C.c_int(0)
    rpp = rp.ctypes.data_as(C.c_void_p)
    check(lib.gdmcf_spmm_plan(rpp, n_rows, chunk, None, 0, None, 0, C.byref(ni), C.byref(nl), C.byref(ns)), "spmm_plan(query)")
    items = np.empty((max(ni.value, 1), 4), dtype=np.int32)
    longs = np.empty((max(nl.value, 1), 3), dtype=np.int32)
    check(lib.gdmcf_spmm_plan(rpp, n_rows, chunk, items.ctypes.data_as(C.c_void_p), ni.value,
                              longs.ctypes.data_as(C.c_void_p), nl.value, C.byref(ni), C.byref(nl), C.byref(ns)), "spmm_plan")
    return SpmmPlan(torch.from_numpy(items).to(device), torch.from_numpy(longs).to(device), ni.value, nl.value,
                    ns.value, n_rows, chunk)  # device="cpu" keeps the plan on the host (lightgcn_plan_bf16 post-processes it)


def spmm_csr(plan: SpmmPlan, col, val, X, Z=None, alpha: float = 1.0, beta: float = 0.0, out=None, scratch=None):
    require_cuda(col, val, X, Z)
    n_rows, d = plan.n_rows, X.shape[1]
    assert X.dtype == torch.float32 and X.is_contiguous()
    if out is None:
        out = torch.empty(n_rows, d, dtype=torch.float32, device=X.device)
    if scratch is None and plan.n_slots > 0:
        scratch = torch.empty(plan.n_slots, d, dtype=torch.float32, device=X.device)
    check(load().gdmcf_spmm_csr_f32(ptr(col), ptr(val), ptr(plan.items), plan.n_items, ptr(plan.long_rows), plan.n_long,
                                    ptr(X), ptr(Z), ptr(out), ptr(scratch), n_rows, d, alpha, beta, stream()), "spmm_csr_f32")
    return out


def lightgcn_propagate(plan: SpmmPlan, col, val, E0, n_layers: int, out=None, work=None, dinv=None):
    """mean_{k<=K} A^k E0 (lightGCN.py:180-194). `work` = (tmp0, tmp1, scratch[, u0]) to avoid reallocation.
    dinv (= D^-1/2 of the binary adjacency, kernels.norm_adj_dinv): use the separable-normalisation form, which needs no
    values (`val` is ignored) — same result up to fp32 rounding."""
    require_cuda(col, val, E0, dinv)
    n, d = E0.shape
    assert n == plan.n_rows and E0.dtype == torch.float32 and E0.is_contiguous()
    if out is None:
        out = torch.empty_like(E0)
    if work is None:
        work = (torch.empty_like(E0), torch.empty_like(E0),
                torch.empty(max(plan.n_slots, 1), d, dtype=torch.float32, device=E0.device))
    if dinv is not None:
        if len(work) < 4 or any(w.shape[0] != n + 1 for w in (work[0], work[1], work[3])):
            work = lightgcn_sym_work(plan, E0)
        tmp0, tmp1, scratch, u0 = work
        check(load().gdmcf_lightgcn_propagate_sym_f32(ptr(col), ptr(dinv), ptr(plan.items), plan.n_items, ptr(plan.long_rows),
                                                      plan.n_long, ptr(E0), ptr(u0), ptr(tmp0), ptr(tmp1), ptr(out),
                                                      ptr(scratch), n, d, n_layers, stream()), "lightgcn_propagate_sym_f32")
        return out
    tmp0, tmp1, scratch = work[:3]
    check(load().gdmcf_lightgcn_propagate_f32(ptr(col), ptr(val), ptr(plan.items), plan.n_items, ptr(plan.long_rows),
                                              plan.n_long, ptr(E0), ptr(tmp0), ptr(tmp1), ptr(out), ptr(scratch), n, d,
                                              n_layers, stream()), "lightgcn_propagate_f32")
    return out


def lightgcn_sym_work(plan: SpmmPlan, E0):
    """(tmp0, tmp1, scratch, u0) for lightgcn_propagate(..., dinv=...): [n + 1, d] buffers whose last row is zero."""
    n, d = E0.shape
    mk = lambda: torch.zeros(n + 1, d, dtype=torch.float32, device=E0.device)  # noqa: E731
    return (mk(), mk(), torch.empty(max(plan.n_slots, 1), d, dtype=torch.float32, device=E0.device), mk())


@dataclass
class LightgcnBf16Plan:
    """Device-side work description of gdmcf_lightgcn_propagate_bf16 (see lightgcn_plan_bf16)."""
    col: torch.Tensor         # int32 [nnz] neighbour lists, hot-first, hot neighbours = 0x80000000 | slot
    items: torch.Tensor       # int32 [n_items, 4]
    mids: torch.Tensor        # int32 [n_items] end of each item's hot prefix
    long_rows: torch.Tensor   # int32 [n_long, 3]
    hot_rows: torch.Tensor    # int32 [n_hot]
    n_items: int
    n_pieces: int             # the first n_pieces items are pieces of hub rows
    warp_ptr: torch.Tensor    # int32 [n_slots + 1] whole-row items of warp slot w: items[n_pieces + warp_ptr[w] : n_pieces + warp_ptr[w + 1]]
    n_slots: int
    n_long: int
    n_hot: int
    n_rows: int
    u0: torch.Tensor
    u1: torch.Tensor
    scratch: torch.Tensor
    sync: torch.Tensor


def lightgcn_plan_bf16(rowptr, col, chunk: int = 128, device="cuda") -> LightgcnBf16Plan:
    """Host-side plan: work items (rows / hub-row pieces), the hot set (most frequently gathered rows, staged in shared
    memory by the kernel) and the hot-first reordering of every row's neighbour list."""
    lib = load()
    rp = np.ascontiguousarray(rowptr.cpu().numpy() if isinstance(rowptr, torch.Tensor) else rowptr, dtype=np.int64)
    cl = np.ascontiguousarray(col.cpu().numpy() if isinstance(col, torch.Tensor) else col, dtype=np.int64)
    n = rp.shape[0] - 1
    base = spmm_plan(rp.astype(np.int32), chunk=chunk, device="cpu")
    items = base.items.numpy().copy()[: base.n_items]
    longs = base.long_rows.numpy().copy()[: max(base.n_long, 0)]
    if base.n_long:
        li_of_row = np.full(n, -1, dtype=np.int64)
        li_of_row[longs[:, 0]] = np.arange(base.n_long)
        piece = items[:, 3] >= 0
        items[piece, 0] = -(li_of_row[items[piece, 0]] + 1)
    cnt = np.bincount(cl, minlength=n)
    n_hot = int(min(lib.gdmcf_lightgcn_hot_rows(), np.count_nonzero(cnt)))
    hot = np.argsort(-cnt, kind="stable")[:n_hot]
    slot_of = np.full(n, -1, dtype=np.int64)
    slot_of[hot] = np.arange(n_hot)
    s = slot_of[cl]
    is_hot = s >= 0
    row_of = np.repeat(np.arange(n), np.diff(rp))
    order = np.lexsort((cl, ~is_hot, row_of))           # by row, hot neighbours first, then ascending id
    enc = np.where(is_hot, s | 0x80000000, cl).astype(np.uint32)[order].view(np.int32)
    hot_prefix = np.concatenate([[0], np.cumsum(is_hot[order])])
    mids = items[:, 1] + (hot_prefix[items[:, 2]] - hot_prefix[items[:, 1]]) if len(items) else np.zeros(1)
    n_pieces = int((items[:, 3] >= 0).sum()) if len(items) else 0
    # Static, cost-balanced dealing. Hub pieces go round-robin over the warp slots (one per warp at a time, kernel part
    # (a)); whole rows are walked two per warp (part (b)), so they are paired by estimated cost, heaviest first, and every
    # pair goes to the warp slot with the least work so far (longest-processing-time greedy, seeded with the slot's pieces).
    # Cost unit = one gather trip of a warp (16 neighbour rows from L2; staged rows ~3x cheaper; +1 per item for the
    # descriptor / epilogue). The grid barrier at the end of a layer waits for the slowest warp, so this is the layer time.
    n_slots = 32 * max(int(lib.gdmcf_num_sms()), 1)
    load_of = np.zeros(n_slots)
    if n_pieces:
        pc = np.ceil((items[:n_pieces, 2] - mids[:n_pieces]) / 16) + 0.35 * np.ceil((mids[:n_pieces] - items[:n_pieces, 1]) / 16) + 1.5
        np.add.at(load_of, np.arange(n_pieces) % n_slots, pc)
    rows = np.arange(n_pieces, len(items))
    warp_ptr = np.zeros(n_slots + 1, dtype=np.int64)
    if len(rows):
        trips = np.ceil((items[rows, 2] - mids[rows]) / 8) + 0.35 * np.ceil((mids[rows] - items[rows, 1]) / 8)
        by_cost = rows[np.argsort(-trips, kind="stable")]
        pair_cost = trips[by_cost - n_pieces][0::2] + 1.0   # the heavier row of each pair sets the trip count
        import heapq
        heap = [(float(l), w) for w, l in enumerate(load_of)]
        heapq.heapify(heap)
        slot_of_pair = np.empty(len(pair_cost), dtype=np.int64)
        for j, c in enumerate(pair_cost.tolist()):
            l, w = heap[0]
            slot_of_pair[j] = w
            heapq.heapreplace(heap, (l + c, w))
        slot_of_row = np.repeat(slot_of_pair, 2)[: len(by_cost)]
        order_r = np.argsort(slot_of_row, kind="stable")  # by slot, heaviest pair first, pairs kept adjacent
        perm = by_cost[order_r]
        warp_ptr[1:] = np.cumsum(np.bincount(slot_of_row, minlength=n_slots))
        items[n_pieces:], mids[n_pieces:] = items[perm], mids[perm]
    i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)  # noqa: E731
    mk = lambda: torch.zeros(n + 1, 64, dtype=torch.bfloat16, device=device)  # noqa: E731
    return LightgcnBf16Plan(torch.from_numpy(enc.copy()).to(device), i32(items if len(items) else np.zeros((1, 4))), i32(mids),
                            i32(longs if len(longs) else np.zeros((1, 3))), i32(hot if n_hot else np.zeros(1)),
                            base.n_items, n_pieces, i32(warp_ptr), n_slots, base.n_long, n_hot, n, mk(), mk(),
                            torch.empty(max(base.n_slots, 1), 64, dtype=torch.float32, device=device),
                            torch.zeros(33 + max(base.n_long, 1) + 32 + 2 * 256, dtype=torch.int32, device=device))


def lightgcn_propagate_bf16(plan: LightgcnBf16Plan, dinv, E0, n_layers: int, out=None):
    """mean_{k<=K} A~^k E0 with bf16 iterated tables, one persistent launch (gdmcf_lightgcn_propagate_bf16)."""
    require_cuda(dinv, E0, plan.col)
    n, d = E0.shape
    assert n == plan.n_rows and d == 64 and E0.dtype == torch.float32 and E0.is_contiguous()
    if out is None:
        out = torch.empty_like(E0)
    check(load().gdmcf_lightgcn_propagate_bf16(ptr(plan.col), ptr(plan.items), ptr(plan.mids), plan.n_items, plan.n_pieces,
                                               ptr(plan.warp_ptr), plan.n_slots,
                                               ptr(plan.long_rows), plan.n_long,
                                               ptr(plan.hot_rows), plan.n_hot, ptr(dinv), ptr(E0), ptr(plan.u0), ptr(plan.u1),
                                               ptr(out), ptr(plan.scratch), ptr(plan.sync), n, d, n_layers, stream()),
          "lightgcn_propagate_bf16")
    return out


def norm_adj_dinv(r_rowptr, rt_rowptr, n_users: int, n_items: int) -> torch.Tensor:
    """dinv[r] = (deg_r + 1e-9)^-1/2 over users then items (lightGCN.py:160-166)."""
    require_cuda(r_rowptr, rt_rowptr)
    dinv = torch.empty(n_users + n_items, dtype=torch.float32, device=r_rowptr.device)
    check(load().gdmcf_norm_adj_dinv(ptr(r_rowptr), ptr(rt_rowptr), n_users, n_items, ptr(dinv), stream()), "norm_adj_dinv")
    return dinv


def build_norm_adj(r_rowptr, r_col, rt_rowptr, rt_col, n_users: int, n_items: int):
    require_cuda(r_rowptr, r_col, rt_rowptr, rt_col)
    nnz = r_col.numel()
    dev = r_col.device
    rowptr = torch.empty(n_users + n_items + 1, dtype=torch.int32, device=dev)
    col = torch.empty(2 * nnz, dtype=torch.int32, device=dev)
    val = torch.empty(2 * nnz, dtype=torch.float32, device=dev)
    check(load().gdmcf_build_norm_adj(ptr(r_rowptr), ptr(r_col), ptr(rt_rowptr), ptr(rt_col), n_users, n_items,
                                      ptr(rowptr), ptr(col), ptr(val), stream()), "build_norm_adj")
    return rowptr, col, val


# ----------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------
_ws_cache: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


_sm_limit = 0


def gemm_set_sm_limit(sms: int) -> None:
    """Upper bound on the SMs the contraction kernels occupy from now on (0 = all). See gdmcf_gemm_set_sm_limit."""
    global _sm_limit
    check(load().gdmcf_gemm_set_sm_limit(int(sms)), "gemm_set_sm_limit")
    _sm_limit = int(sms)


def gemm(a: Sequence[torch.Tensor], b: Sequence[torch.Tensor], m: int, n: int, k: Sequence[int], *,
         mode: int = EPI_STORE, act: int = ACT_NONE, alpha: float = 1.0, out_f32=None, out_bf16=None, out_bf16_lo=None,
         bias=None, ld_bias: int = 0, row_t=None, t_const: int = 0, row_scale=None, col_scale=None, c1=None, c2=None,
         xt=None, splits: Optional[int] = None) -> None:
    """C[m,n] = sum_s a[s][m,k[s]] @ b[s][n,k[s]]^T with a fused epilogue (gdmcf_gemm_bf16_tn).

    a[s]/b[s]: 2-D bf16 tensors whose stride(0) is the leading dimension (stride(1) == 1)."""
    lib = load()
    g = GemmDesc()
    assert 1 <= len(a) == len(b) == len(k) <= _lib.MAX_SEG
    for s, (ta, tb, ks) in enumerate(zip(a, b, k)):
        require_cuda(ta, tb)
        assert ta.dtype == torch.bfloat16 and tb.dtype == torch.bfloat16 and ta.stride(1) == 1 and tb.stride(1) == 1
        g.a[s], g.b[s] = ta.data_ptr(), tb.data_ptr()
        g.lda[s], g.ldb[s] = ta.stride(0), tb.stride(0)
        g.k[s] = ks
    g.n_seg, g.m, g.n = len(a), m, n
    e = Epilogue()
    e.mode, e.act, e.alpha, e.t_const = mode, act, alpha, t_const
    for t in (out_f32, out_bf16, out_bf16_lo, xt):
        assert t is None or t.stride(1) == 1
    e.out_f32, e.ld_f32 = ptr(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    e.out_bf16, e.ld_bf16 = ptr(out_bf16), (out_bf16.stride(0) if out_bf16 is not None else 0)
    e.out_bf16_lo = ptr(out_bf16_lo)
    if out_bf16_lo is not None:
        assert out_bf16 is not None and out_bf16_lo.stride(0) == out_bf16.stride(0)
    e.bias, e.ld_bias = ptr(bias), ld_bias
    e.row_t = ptr(row_t)
    e.row_scale, e.col_scale = ptr(row_scale), ptr(col_scale)
    e.c1, e.c2 = ptr(c1), ptr(c2)
    e.xt, e.ld_xt = ptr(xt), (xt.stride(0) if xt is not None else 0)
    if splits is None:
        splits = lib.gdmcf_gemm_auto_splits(m, n, int(sum(round_up(x, 64) for x in k)))
    ws, ws_bytes = None, 0
    if splits > 1:
        ws_bytes = lib.gdmcf_gemm_workspace_bytes(m, n, splits)
        ws = _workspace(ws_bytes, a[0].device)
    check(lib.gdmcf_gemm_bf16_tn(C.byref(g), C.byref(e), splits, ptr(ws), ws_bytes, stream()), "gemm_bf16_tn")


_tower_ws: dict = {}


def user_tower(hc: Bf16Mat, hc_f32, w1: Bf16Mat, b1, w2: Bf16Mat, b2, sumw, rows: int, *, out: Bf16Mat, inv_u,
               g1_f32=None, g2_f32=None, hcp_f32=None, max_ctas: Optional[int] = None) -> None:
    """conv1 + relu + conv2 + sumW mix + user norms of the GDMCF tower in one launch (gdmcf_user_tower, bf16 mode).
    max_ctas defaults to the SM limit currently imposed on the contractions (gemm_set_sm_limit)."""
    if max_ctas is None:
        max_ctas = _sm_limit
    require_cuda(hc.hi, hc_f32, w1.hi, b1, w2.hi, b2, sumw, out.hi, inv_u, g1_f32, g2_f32, hcp_f32)
    k1, hidden, n = hc.cols, w1.rows, w2.rows
    assert w1.cols == k1 and w2.cols == hidden and out.cols == n and hc_f32.stride(1) == 1
    lib = load()
    key = (str(hc.hi.device), torch.cuda.current_stream().cuda_stream, rows, k1, hidden, n)
    ws = _tower_ws.get(key)
    if ws is None:
        nbytes = lib.gdmcf_user_tower_workspace_bytes(rows, k1, hidden, n)
        ws = (torch.empty(nbytes, dtype=torch.uint8, device=hc.hi.device), torch.zeros(16, dtype=torch.int32, device=hc.hi.device))
        _tower_ws[key] = ws
    check(lib.gdmcf_user_tower(ptr(hc.hi), hc.ld, ptr(hc_f32), hc_f32.stride(0), ptr(w1.hi), w1.ld, ptr(b1), ptr(w2.hi), w2.ld,
                               ptr(b2), ptr(sumw), rows, k1, hidden, n, ptr(out.hi), out.ld, ptr(inv_u), ptr(g1_f32), ptr(g2_f32),
                               g2_f32.stride(0) if g2_f32 is not None else 0, ptr(hcp_f32),
                               hcp_f32.stride(0) if hcp_f32 is not None else 0, ptr(ws[0]), ws[0].numel(), ptr(ws[1]), max_ctas,
                               stream()), "user_tower")


# ----------------------------------------------------------------------------------------------
# elementwise / layout
# ----------------------------------------------------------------------------------------------
def cast_bf16(x: torch.Tensor, with_lo: bool = False, out: Optional[Bf16Mat] = None) -> Bf16Mat:
    require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        out = Bf16Mat.empty(rows, cols, x.device, with_lo, zero=False)
    check(load().gdmcf_cast_bf16(ptr(x), x.stride(0), ptr(out.hi), ptr(out.lo), out.ld, rows, cols, stream()), "cast_bf16")
    return out


def cast_bf16_transpose(x: torch.Tensor, with_lo: bool = False, out: Optional[Bf16Mat] = None) -> Bf16Mat:
    """Returns x^T as a bf16 operand [cols, rows]."""
    require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        out = Bf16Mat.empty(cols, rows, x.device, with_lo, zero=False)
    check(load().gdmcf_cast_bf16_transpose(ptr(x), x.stride(0), ptr(out.hi), ptr(out.lo), out.ld, rows, cols, stream()),
          "cast_bf16_transpose")
    return out


def scale_cols_cast(x: torch.Tensor, cols: int, col_scale: torch.Tensor, with_lo: bool = False, out: Optional[Bf16Mat] = None) -> Bf16Mat:
    """bf16 operand of x[:, :cols] * col_scale[None, :] (gdmcf_scale_cols_cast)."""
    require_cuda(x, col_scale)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and col_scale.numel() >= cols
    rows = x.shape[0]
    if out is None:
        out = Bf16Mat.empty(rows, cols, x.device, with_lo, zero=False)
    check(load().gdmcf_scale_cols_cast(ptr(x), x.stride(0), ptr(col_scale), ptr(out.hi), ptr(out.lo), out.ld, rows, cols, stream()),
          "scale_cols_cast")
    return out


def densify_rows(rowptr, col, users, n_rows: int, n_items: int, out_f32=None, out_bf16=None) -> None:
    require_cuda(rowptr, col, users, out_f32, out_bf16)
    check(load().gdmcf_densify_rows(ptr(rowptr), ptr(col), ptr(users), n_rows, n_items, ptr(out_f32),
                                    out_f32.stride(0) if out_f32 is not None else 0, ptr(out_bf16),
                                    out_bf16.stride(0) if out_bf16 is not None else 0, stream()), "densify_rows")


def qsample_dropout(x0, rows: int, cols: int, a_out: Bf16Mat, *, row_t=None, t_const: int = 0, sqrt_ab=None,
                    sqrt_1mab=None, noise=None, keep=None, dropout_p: float = 0.0, seed: int = 0, offset: int = 0,
                    epoch=None, xt_out=None) -> None:
    require_cuda(x0, a_out.hi, row_t, sqrt_ab, sqrt_1mab, noise, keep, xt_out)
    if noise is not None:
        assert noise.is_contiguous() and noise.shape == (rows, cols) and noise.dtype == torch.float32
    if keep is not None:
        assert keep.is_contiguous() and keep.shape == (rows, cols) and keep.dtype == torch.uint8
    check(load().gdmcf_qsample_dropout(ptr(x0), x0.stride(0), ptr(row_t), t_const, ptr(sqrt_ab), ptr(sqrt_1mab), ptr(noise),
                                       ptr(keep), dropout_p, seed, offset, ptr(epoch), ptr(xt_out),
                                       xt_out.stride(0) if xt_out is not None else 0, ptr(a_out.hi), ptr(a_out.lo),
                                       a_out.ld, rows, cols, stream()), "qsample_dropout")


def onehot_noise(x0, rows: int, cols: int, out: torch.Tensor, *, ts=None, discrete: float = 0.9995,
                 dropout_p: float = 0.0, u_keep=None, u_drop=None, seed: int = 0, offset: int = 0, epoch=None) -> None:
    require_cuda(x0, out, ts, u_keep, u_drop)
    assert out.dtype == torch.bfloat16 and out.stride(1) == 1
    check(load().gdmcf_onehot_noise(ptr(x0), x0.stride(0), ptr(ts), discrete, dropout_p, ptr(u_keep), ptr(u_drop), seed,
                                    offset, ptr(epoch), ptr(out), out.stride(0), rows, cols, stream()), "onehot_noise")


def graph_noise_step(state: torch.Tensor, t: int, batch: int, *, deg_frac=None, discrete: float = 0.9995, user_guided: bool = True,
                     seed: int = 0, offset: int = 0, epoch=None, u_entry=None, u_user=None) -> None:
    """One reverse step of p_sample's random graph bookkeeping on a uint8 edge matrix [rows, cols] (gdmcf_graph_noise_step)."""
    require_cuda(state, deg_frac, epoch, u_entry, u_user)
    assert state.dtype == torch.uint8 and state.dim() == 2 and state.stride(1) == 1
    rows, cols = state.shape
    check(load().gdmcf_graph_noise_step(ptr(state), state.stride(0), ptr(deg_frac), int(t), int(batch), discrete, int(user_guided),
                                        seed, offset, ptr(epoch), ptr(u_entry), ptr(u_user), rows, cols, stream()), "graph_noise_step")


def onehot_tables(w2: torch.Tensor, d: int, n_items: int, out=None):
    """out: optional (base, delta) pair from an earlier call, refreshed in place (stable addresses for CUDA graphs)."""
    require_cuda(w2)
    assert w2.dtype == torch.float32 and w2.stride(1) == 1
    ld_delta = round_up(d, 4)
    if out is not None and out[0].shape == (d,) and out[1].shape == (n_items, ld_delta) and out[0].device == w2.device:
        base, delta = out
    else:
        base = torch.empty(d, dtype=torch.float32, device=w2.device)
        delta = torch.empty(n_items, ld_delta, dtype=torch.float32, device=w2.device)
    # in_layers2.0.weight is [d, 2I + e]: its row stride is odd when e is; the kernel needs float2-aligned
    # (2i, 2i+1) pairs, so stage the [d, 2I] block with an even leading dimension first when required.
    if w2.stride(0) % 2 or w2.data_ptr() % 8:
        w2 = w2[:, : 2 * n_items].contiguous()
    check(load().gdmcf_onehot_tables(ptr(w2), w2.stride(0), d, n_items, ptr(base), ptr(delta), ld_delta, stream()),
          "onehot_tables")
    return base, delta


_gather_ws: dict = {}


def encode_onehot_gather(rowptr, col, users, n_rows: int, base, delta, d: int, out: torch.Tensor) -> None:
    require_cuda(rowptr, col, users, base, delta, out)
    lib = load()
    key = (str(out.device), torch.cuda.current_stream().cuda_stream, n_rows, d)
    ws = _gather_ws.get(key)
    if ws is None:  # partial sums of the heavy users' row slices + per-row completion counters (left zeroed by the kernel)
        ws = (torch.empty(lib.gdmcf_encode_onehot_gather_workspace_bytes(n_rows, d) // 4, dtype=torch.float32, device=out.device),
              torch.zeros(n_rows, dtype=torch.int32, device=out.device))
        _gather_ws[key] = ws
    check(lib.gdmcf_encode_onehot_gather(ptr(rowptr), ptr(col), ptr(users), n_rows, ptr(base), ptr(delta), delta.stride(0), d,
                                         ptr(out), out.stride(0), ptr(ws[0]), ws[0].numel() * 4, ptr(ws[1]), stream()),
          "encode_onehot_gather")


def time_bias_table(w_emb, b_emb, w_layer: torch.Tensor, n_in: int, bias, T: int, out=None):
    """[T, d] table of first-layer biases; w_layer is the layer's full fp32 weight [d, n_in + e].
    out: optional (table, emb_table) pair from an earlier call, refreshed in place."""
    require_cuda(w_emb, b_emb, w_layer, bias)
    d, e = w_layer.shape[0], w_layer.shape[1] - n_in
    assert w_emb.shape == (e, e) and w_emb.is_contiguous() and w_layer.stride(1) == 1
    if out is not None and out[0].shape == (T, d) and out[1].shape == (T, e) and out[0].device == w_layer.device:
        out, emb_table = out
    else:
        emb_table = torch.empty(T, e, dtype=torch.float32, device=w_layer.device)
        out = torch.empty(T, d, dtype=torch.float32, device=w_layer.device)
    w_time = w_layer[:, n_in:]
    check(load().gdmcf_time_bias_table(ptr(w_emb), ptr(b_emb), w_time.data_ptr(), w_layer.stride(0), ptr(bias), T, e, d,
                                       ptr(emb_table), ptr(out), d, stream()), "time_bias_table")
    return out, emb_table


def bias_act_rows(x, rows: int, cols: int, *, bias=None, ld_bias: int = 0, row_t=None, t_const: int = 0,
                  act: int = ACT_NONE, out_f32=None, out_bf16=None, out_bf16_lo=None) -> None:
    require_cuda(x, bias, row_t, out_f32, out_bf16, out_bf16_lo)
    check(load().gdmcf_bias_act_rows(ptr(x), x.stride(0), ptr(bias), ld_bias, ptr(row_t), t_const, act, ptr(out_f32),
                                     out_f32.stride(0) if out_f32 is not None else 0, ptr(out_bf16), ptr(out_bf16_lo),
                                     out_bf16.stride(0) if out_bf16 is not None else 0, rows, cols, stream()), "bias_act_rows")


def gather_rows(table, idx, rows: int, cols: int, *, out_f32=None, out_bf16=None, out_bf16_lo=None) -> None:
    require_cuda(table, idx, out_f32, out_bf16, out_bf16_lo)
    assert idx.dtype == torch.int32
    check(load().gdmcf_gather_rows(ptr(table), table.stride(0), ptr(idx), ptr(out_f32),
                                   out_f32.stride(0) if out_f32 is not None else 0, ptr(out_bf16), ptr(out_bf16_lo),
                                   out_bf16.stride(0) if out_bf16 is not None else 0, rows, cols, stream()), "gather_rows")


def sgemm_small(A, B, C, m: int, n: int, k: int, *, trans_a=False, trans_b=False, alpha=1.0, beta=0.0) -> None:
    require_cuda(A, B, C)
    assert A.dtype == B.dtype == C.dtype == torch.float32 and A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
    check(load().gdmcf_sgemm_small(ptr(A), A.stride(0), int(trans_a), ptr(B), B.stride(0), int(trans_b), ptr(C), C.stride(0),
                                   m, n, k, alpha, beta, stream()), "sgemm_small")


def colsum_f32(x, rows: int, cols: int, out=None) -> torch.Tensor:
    require_cuda(x)
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=x.device)
    check(load().gdmcf_colsum_f32(ptr(x), x.stride(0), rows, cols, ptr(out), stream()), "colsum_f32")
    return out


def mix_rownorm(hc, rows: int, cols: int, *, g=None, sumw=None, out_f32=None, out: Optional[Bf16Mat] = None,
                inv_norm=None) -> None:
    require_cuda(hc, g, sumw, out_f32, inv_norm)
    check(load().gdmcf_mix_rownorm(ptr(hc), hc.stride(0), ptr(g), g.stride(0) if g is not None else 0, ptr(sumw),
                                   ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
                                   ptr(out.hi) if out is not None else None, ptr(out.lo) if out is not None else None,
                                   out.ld if out is not None else 0, ptr(inv_norm), rows, cols, stream()), "mix_rownorm")


def row_inv_norm(x: torch.Tensor, out=None) -> torch.Tensor:
    require_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    if out is None:
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    check(load().gdmcf_row_inv_norm(ptr(x), x.stride(0), ptr(out), x.shape[0], x.shape[1], stream()), "row_inv_norm")
    return out


# ----------------------------------------------------------------------------------------------
# ranking
# ----------------------------------------------------------------------------------------------
def mask_topk(scores, n_rows: int, n_items: int, k: int, *, users=None, hist=None, hist2=None, with_values=False):
    """hist / hist2: (rowptr, col) int32 device CSR pairs. Returns idx int32 [n_rows, k] (and values)."""
    require_cuda(scores, users)
    assert scores.dtype == torch.float32 and scores.stride(1) == 1
    idx = torch.empty(n_rows, k, dtype=torch.int32, device=scores.device)
    val = torch.empty(n_rows, k, dtype=torch.float32, device=scores.device) if with_values else None
    h1 = hist if hist is not None else (None, None)
    h2 = hist2 if hist2 is not None else (None, None)
    check(load().gdmcf_mask_topk(ptr(scores), scores.stride(0), n_rows, n_items, ptr(users), ptr(h1[0]), ptr(h1[1]),
                                 ptr(h2[0]), ptr(h2[1]), k, ptr(idx), ptr(val), stream()), "mask_topk")
    return (idx, val) if with_values else idx


def topn_metrics(topk_idx, users, gt_rowptr, gt_col, topn_dev, n_topn: int) -> torch.Tensor:
    require_cuda(topk_idx, users, gt_rowptr, gt_col, topn_dev)
    n_rows = topk_idx.shape[0]
    stats = torch.empty(n_rows, n_topn, 4, dtype=torch.float64, device=topk_idx.device)
    check(load().gdmcf_topn_metrics(ptr(topk_idx), topk_idx.stride(0), n_rows, ptr(users), ptr(gt_rowptr), ptr(gt_col),
                                    ptr(topn_dev), n_topn, ptr(stats), stream()), "topn_metrics")
    return stats


def colsum_f64(x: torch.Tensor) -> torch.Tensor:
    require_cuda(x)
    x2 = x.reshape(x.shape[0], -1)
    assert x2.dtype == torch.float64 and x2.is_contiguous()
    out = torch.empty(x2.shape[1], dtype=torch.float64, device=x.device)
    check(load().gdmcf_colsum_f64(ptr(x2), x2.shape[0], x2.shape[1], ptr(out), stream()), "colsum_f64")
    return out


# ----------------------------------------------------------------------------------------------
# training
# ----------------------------------------------------------------------------------------------
def mse_rows(out, x0, rows: int, cols: int) -> torch.Tensor:
    require_cuda(out, x0)
    mse = torch.empty(rows, dtype=torch.float32, device=out.device)
    check(load().gdmcf_mse_rows(ptr(out), out.stride(0), ptr(x0), x0.stride(0), rows, cols, ptr(mse), stream()), "mse_rows")
    return mse


def adamw_fused(p, g, m, v, *, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                weight_decay: float = 0.0, step: int = 1, step_dev=None, grad_scale: float = 1.0) -> None:
    require_cuda(p, g, m, v)
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    check(load().gdmcf_adamw_fused(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                   ptr(step_dev), grad_scale, stream()), "adamw_fused")


def adamw_rows_lazy(p, m, v, last_step, *, idx=None, grad_rows=None, n_sel: int = 0, lr: float, beta1: float = 0.9,
                    beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0, step: int = 1, step_dev=None,
                    grad_scale: float = 1.0) -> None:
    """Row-sparse AdamW with exact catch-up of the skipped steps (gdmcf_adamw_rows_lazy). idx=None: flush all rows."""
    require_cuda(p, m, v, last_step, idx, grad_rows)
    assert p.dim() == 2 and p.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    assert last_step.dtype == torch.int32 and last_step.numel() == p.shape[0]
    ld_g = 0
    if idx is not None:
        assert idx.dtype == torch.int32 and idx.is_contiguous() and n_sel > 0
    if grad_rows is not None:
        assert idx is not None and grad_rows.stride(1) == 1 and grad_rows.shape[1] == p.shape[1]
        ld_g = grad_rows.stride(0)
    check(load().gdmcf_adamw_rows_lazy(ptr(p), ptr(m), ptr(v), ptr(last_step), ptr(idx), ptr(grad_rows), ld_g, n_sel, p.shape[0],
                                       p.shape[1], lr, beta1, beta2, eps, weight_decay, step, ptr(step_dev), grad_scale, stream()),
          "adamw_rows_lazy")


def adamw_partitioned(p, g, m, v, *, n_ctas: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                      weight_decay: float = 0.0, step: int = 1, step_dev=None, grad_scale: float = 1.0, row_coef=None) -> None:
    """AdamW (no derived tensors) confined to n_ctas SMs (gdmcf_adamw_partitioned): runs next to SM-limited contractions
    on another stream. p/m/v contiguous; g may be a [rows, cols] view with a padded leading dimension."""
    require_cuda(p, g, m, v, row_coef)
    assert p.is_contiguous() and m.is_contiguous() and v.is_contiguous() and g.shape == p.shape
    if p.dim() == 2:
        rows, cols = p.shape
        assert g.stride(1) == 1
        ld_g = g.stride(0)
    else:
        rows, cols, ld_g = 1, p.numel(), p.numel()
        assert g.is_contiguous()
    check(load().gdmcf_adamw_partitioned(ptr(p), ptr(g), ld_g, ptr(m), ptr(v), rows, cols, lr, beta1, beta2, eps, weight_decay,
                                         step, ptr(step_dev), grad_scale, ptr(row_coef), int(n_ctas), stream()),
          "adamw_partitioned")


def adamw_refresh(p, g, m, v, *, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                  weight_decay: float = 0.0, step: int = 1, step_dev=None, grad_scale: float = 1.0, cols_used: int = 0,
                  op: Optional[Bf16Mat] = None, op_t: Optional[Bf16Mat] = None, inv_norm=None, delta=None, base=None,
                  tcols=None, rowpart=None, row_coef=None) -> None:
    """AdamW on a 2-D weight fused with the refresh of its derived tensors (gdmcf_adamw_refresh). g may be a
    [rows, cols] view with a padded leading dimension."""
    require_cuda(p, g, m, v, inv_norm, delta, base, tcols, rowpart, row_coef)
    assert p.dim() == 2 and p.is_contiguous() and m.is_contiguous() and v.is_contiguous() and g.shape == p.shape and g.stride(1) == 1
    rows, cols = p.shape
    r = _lib.Refresh()
    r.cols_used = cols_used
    if op is not None:
        r.hi, r.lo, r.ld_hi = ptr(op.hi), ptr(op.lo), op.ld
    if op_t is not None:
        r.t_hi, r.t_lo, r.ld_t = ptr(op_t.hi), ptr(op_t.lo), op_t.ld
    r.inv_norm = ptr(inv_norm)
    if delta is not None:
        r.delta, r.ld_delta, r.base = ptr(delta), delta.stride(0), ptr(base)
    if tcols is not None:
        assert tcols.is_contiguous()
        r.tcols, r.n_tcols = ptr(tcols), tcols.shape[1]
    if row_coef is not None:
        assert row_coef.dtype == torch.float32 and row_coef.is_contiguous() and row_coef.numel() == rows
        r.row_coef = ptr(row_coef)
    if inv_norm is not None or delta is not None:
        assert rowpart is not None and rowpart.numel() >= adamw_refresh_splits(rows, cols) * rows
        r.rowpart = ptr(rowpart)
    check(load().gdmcf_adamw_refresh(ptr(p), ptr(g), g.stride(0), ptr(m), ptr(v), rows, cols, lr, beta1, beta2, eps,
                                     weight_decay, step, ptr(step_dev), grad_scale, C.byref(r), stream()), "adamw_refresh")


def refresh_derived(p, *, cols_used: int = 0, op: Optional[Bf16Mat] = None, op_t: Optional[Bf16Mat] = None, inv_norm=None,
                    delta=None, base=None, tcols=None, rowpart=None) -> None:
    """The derived tensors of a 2-D weight, computed from its current value by the same kernel that refreshes them during
    training (gdmcf_adamw_refresh with g == NULL), so that lazily built and in-training values agree bit for bit."""
    require_cuda(p, inv_norm, delta, base, tcols, rowpart)
    assert p.dim() == 2 and p.is_contiguous()
    rows, cols = p.shape
    r = _lib.Refresh()
    r.cols_used = cols_used
    if op is not None:
        r.hi, r.lo, r.ld_hi = ptr(op.hi), ptr(op.lo), op.ld
    if op_t is not None:
        r.t_hi, r.t_lo, r.ld_t = ptr(op_t.hi), ptr(op_t.lo), op_t.ld
    r.inv_norm = ptr(inv_norm)
    if delta is not None:
        r.delta, r.ld_delta, r.base = ptr(delta), delta.stride(0), ptr(base)
    if tcols is not None:
        r.tcols, r.n_tcols = ptr(tcols), tcols.shape[1]
    if inv_norm is not None or delta is not None:
        if rowpart is None:
            rowpart = torch.empty(adamw_refresh_splits(rows, cols) * rows, dtype=torch.float32, device=p.device)
        r.rowpart = ptr(rowpart)
    check(load().gdmcf_adamw_refresh(ptr(p), None, 0, None, None, rows, cols, 0.0, 0.9, 0.999, 1e-8, 0.0, 1, None, 1.0,
                                     C.byref(r), stream()), "adamw_refresh(refresh only)")


def adamw_refresh_splits(rows: int, cols: int) -> int:
    return load().gdmcf_adamw_refresh_splits(rows, cols)


def loss_grad(out, x0, gs, rows: int, cols: int, G: Bf16Mat, *, GT: Optional[Bf16Mat] = None, row_scale=None,
              col_scale=None, with_out: bool = False, colsum=None, rowpart=None) -> None:
    require_cuda(out, x0, gs, G.hi, row_scale, col_scale, colsum, rowpart)
    check(load().gdmcf_loss_grad(ptr(out), out.stride(0), ptr(x0), x0.stride(0), ptr(gs), ptr(row_scale), ptr(col_scale),
                                 int(with_out), ptr(G.hi), ptr(G.lo), G.ld, ptr(GT.hi) if GT is not None else None,
                                 ptr(GT.lo) if GT is not None else None, GT.ld if GT is not None else 0,
                                 ptr(colsum), ptr(rowpart), rows, cols, stream()), "loss_grad")


def transpose_bf16(x: torch.Tensor, rows: int, cols: int, out: torch.Tensor) -> None:
    require_cuda(x, out)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    check(load().gdmcf_transpose_bf16(ptr(x), x.stride(0), ptr(out), out.stride(0), rows, cols, stream()), "transpose_bf16")


EW_RELU_BWD, EW_TANH_BWD, EW_AXPBY = 0, 1, 2


def ew_binary(op: int, a, b, rows: int, cols: int, *, alpha=1.0, beta=1.0, out_f32=None, out_bf16=None, out_bf16_lo=None) -> None:
    require_cuda(a, b, out_f32, out_bf16, out_bf16_lo)
    check(load().gdmcf_ew_binary(op, ptr(a), a.stride(0), ptr(b), b.stride(0), alpha, beta, ptr(out_f32),
                                 out_f32.stride(0) if out_f32 is not None else 0, ptr(out_bf16), ptr(out_bf16_lo),
                                 out_bf16.stride(0) if out_bf16 is not None else 0, rows, cols, stream()), "ew_binary")


def mix_backward(d_hcp, hc, g2, sumw, d_hc, d_g2, dw_rows, rows: int, cols: int) -> None:
    require_cuda(d_hcp, hc, g2, sumw, d_hc, d_g2, dw_rows)
    check(load().gdmcf_mix_backward(ptr(d_hcp), d_hcp.stride(0), ptr(hc), hc.stride(0), ptr(g2), g2.stride(0), ptr(sumw),
                                    ptr(d_hc), d_hc.stride(0), ptr(d_g2), d_g2.stride(0), ptr(dw_rows), rows, cols, stream()),
          "mix_backward")


def ntxent_rows(S, n: int, *, tau=0.1, eps=1e-5, dscale=None, loss_rows=None, dS=None) -> None:
    require_cuda(S, dscale, loss_rows, dS)
    check(load().gdmcf_ntxent_rows(ptr(S), S.stride(0), n, tau, eps, ptr(dscale), ptr(loss_rows), ptr(dS),
                                   dS.stride(0) if dS is not None else 0, stream()), "ntxent_rows")


def scatter_rows_add(v, idx, grad, rows: int, cols: int) -> None:
    require_cuda(v, idx, grad)
    assert idx.dtype == torch.int32
    check(load().gdmcf_scatter_rows_add(ptr(v), v.stride(0), ptr(idx), ptr(grad), grad.stride(0), rows, cols, stream()),
          "scatter_rows_add")


def loss_terms(ts, pt, mse, alphas_cumprod, closs, reweight: bool):
    """(hist_loss f64 [B], loss f64 [B], g_mse f32 [B]) of training_losses in one launch (gdmcf_loss_terms)."""
    require_cuda(ts, pt, mse, alphas_cumprod, closs)
    B = ts.numel()
    assert ts.dtype == torch.int64 and pt.dtype == torch.float64 and mse.dtype == torch.float32 and alphas_cumprod.dtype == torch.float64
    assert closs is None or closs.dtype == torch.float32
    assert ts.is_contiguous() and pt.is_contiguous() and mse.is_contiguous() and alphas_cumprod.is_contiguous()
    hist = torch.empty(B, dtype=torch.float64, device=ts.device)
    loss = torch.empty(B, dtype=torch.float64, device=ts.device)
    g = torch.empty(B, dtype=torch.float32, device=ts.device)
    check(load().gdmcf_loss_terms(ptr(ts), ptr(pt), ptr(mse), ptr(alphas_cumprod), ptr(closs), B, alphas_cumprod.numel(),
                                  int(bool(reweight)), ptr(hist), ptr(loss), ptr(g), stream()), "loss_terms")
    return hist, loss, g


def lt_history_update(ts, loss, lt_history, lt_count) -> None:
    """In-place Lt_history/Lt_count update (gaussian_diffusion.py:935-949). ts int64 [B], loss fp64 [B]."""
    require_cuda(ts, loss, lt_history, lt_count)
    assert ts.dtype == torch.int64 and loss.dtype == torch.float64 and lt_history.dtype == torch.float64 and lt_count.dtype == torch.int64
    assert ts.is_contiguous() and loss.is_contiguous() and lt_history.is_contiguous() and lt_count.is_contiguous()
    check(load().gdmcf_lt_history_update(ptr(ts), ptr(loss), ptr(lt_history), ptr(lt_count), ts.numel(), lt_history.shape[0],
                                         lt_history.shape[1], stream()), "lt_history_update")


def sample_timesteps(lt_history, lt_count, batch: int, *, uniform_prob: float = 0.001, seed: int = 0, offset: int = 0,
                     epoch=None, ts_in=None):
    """Importance timestep sampling on the device (gaussian_diffusion.py:959-986). Returns (ts int64 [B], pt fp64 [B])."""
    require_cuda(lt_history, lt_count, epoch, ts_in)
    assert lt_history.dtype == torch.float64 and lt_count.dtype == torch.int64 and lt_history.is_contiguous()
    dev = lt_history.device
    if ts_in is not None:
        assert ts_in.dtype == torch.int64 and ts_in.is_contiguous() and ts_in.numel() == batch
    ts = ts_in if ts_in is not None else torch.empty(batch, dtype=torch.int64, device=dev)
    pt = torch.empty(batch, dtype=torch.float64, device=dev)
    check(load().gdmcf_sample_timesteps(ptr(lt_history), ptr(lt_count), lt_history.shape[0], lt_history.shape[1], batch,
                                        uniform_prob, seed, offset, ptr(epoch), ptr(ts_in), ptr(ts), ptr(pt), stream()),
          "sample_timesteps")
    return ts, pt


def counter_add(counter: torch.Tensor, inc: int = 1) -> None:
    """counter[0] += inc on the stream (device-resident step / RNG-epoch counters; int64 or uint64 scalar tensor)."""
    require_cuda(counter)
    assert counter.element_size() == 8
    check(load().gdmcf_counter_add(ptr(counter), inc, stream()), "counter_add")
