"""LightGCN propagation — mirror of the reference's lightGCN.py:129-203 (class LightGCN: init_embedding,
get_A_tilda, propagate_through_layers). The rest of that script (ml-100k loading, BPR training, pandas metrics)
is outside the hot path. The normalized adjacency is built on the device from the interaction CSR
(gdmcf_build_norm_adj) instead of the 163 s dok/lil construction, and propagation is the Horner-form CSR SpMM
(gdmcf_lightgcn_propagate_f32) instead of K torch.sparse.mm calls + stack + mean."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn

from . import kernels as K


class LightGCN(nn.Module):
    def __init__(self, data, n_users, n_items, n_layers, latent_dim, device="cuda", chunk=128, precision="fp32"):
        """data: mapping with 'user_id_idx' / 'item_id_idx' columns (DataFrame or dict of arrays), as the reference.
        precision: "fp32" (1e-5 mode, the spmm.cu kernels) or "bf16" (bf16 iterated tables + shared-memory staged hot rows
        in one persistent launch, spmm_bf16.cu; ~1e-3 normwise on the layer mean; latent_dim must be 64)."""
        super().__init__()
        assert precision in ("fp32", "bf16")
        self.precision = precision
        if latent_dim % 64:
            raise NotImplementedError("latent_dim must be a multiple of 64 (the reference uses 64)")
        self.data, self.n_users, self.n_items = data, n_users, n_items
        self.n_layers, self.latent_dim, self.device, self.chunk = n_layers, latent_dim, device, chunk
        self.init_embedding()
        self.norm_adj_csr = self.get_A_tilda()

    def init_embedding(self):
        self.E0 = nn.Embedding(self.n_users + self.n_items, self.latent_dim)
        nn.init.xavier_uniform_(self.E0.weight)
        self.E0.weight = nn.Parameter(self.E0.weight.to(self.device))

    def get_A_tilda(self):
        """lightGCN.py:145-178 -> device CSR (rowptr, col, val) of D^-1/2 [[0,R],[R^T,0]] D^-1/2 + the SpMM plan."""
        u = np.asarray(self.data['user_id_idx'], dtype=np.int64)
        i = np.asarray(self.data['item_id_idx'], dtype=np.int64)
        R = sp.csr_matrix((np.ones(len(u), dtype=np.float32), (u, i)), shape=(self.n_users, self.n_items))
        R.sum_duplicates()
        R.sort_indices()
        RT = R.T.tocsr()
        RT.sort_indices()
        dev = self.device
        t = lambda a: torch.from_numpy(a.astype(np.int32)).to(dev)  # noqa: E731
        r_rp, rt_rp = t(R.indptr), t(RT.indptr)
        rowptr, col, val = K.build_norm_adj(r_rp, t(R.indices), rt_rp, t(RT.indices), self.n_users, self.n_items)
        # D^-1/2 of the interaction graph (pattern only, like build_norm_adj): the propagation uses the separable form
        self.dinv = K.norm_adj_dinv(r_rp, rt_rp, self.n_users, self.n_items)
        self.plan = K.spmm_plan(rowptr.cpu(), chunk=self.chunk, device=dev)
        self.plan16 = K.lightgcn_plan_bf16(rowptr, col, chunk=self.chunk, device=dev) if self.precision == "bf16" else None
        self.norm_adj_mat_sparse_tensor = (rowptr, col, val)  # reference attribute name; CSR triple here
        return rowptr, col, val

    def propagate_through_layers(self):
        """lightGCN.py:180-194: returns (final_user, final_item, initial_user, initial_item). Differentiable w.r.t. E0
        like the reference's torch.sparse.mm chain (its BPR loop trains through it): Â is symmetric, so the backward of
        mean_k Â^k is the same propagation applied to the incoming gradient."""
        E0 = self.E0.weight
        if torch.is_grad_enabled() and E0.requires_grad:
            mean = _Propagate.apply(E0, self)
        else:
            mean = self._propagate(E0.detach())
        final_user, final_item = torch.split(mean, [self.n_users, self.n_items])
        initial_user, initial_item = torch.split(E0, [self.n_users, self.n_items])
        return final_user, final_item, initial_user, initial_item

    def _propagate(self, X: torch.Tensor) -> torch.Tensor:
        _, col, val = self.norm_adj_csr
        if self.plan16 is not None:
            return K.lightgcn_propagate_bf16(self.plan16, self.dinv, X.contiguous(), self.n_layers)
        return K.lightgcn_propagate(self.plan, col, val, X.contiguous(), self.n_layers, dinv=self.dinv)

    def forward(self, users, pos_items, neg_items):
        fu, fi, iu, ii = self.propagate_through_layers()
        return fu[users], fi[pos_items], fi[neg_items], iu[users], ii[pos_items], ii[neg_items]


class _Propagate(torch.autograd.Function):
    """mean_{k<=K} Â^k X with the hand-written SpMM in both directions (no torch.sparse path)."""

    @staticmethod
    def forward(ctx, E0, lg):
        ctx.lg = lg
        return lg._propagate(E0.detach())

    @staticmethod
    def backward(ctx, grad):
        return ctx.lg._propagate(grad), None
