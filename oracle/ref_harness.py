"""TEST INFRASTRUCTURE — imports the *unmodified* reference (/root/reference) on CPU so that golden vectors
can be generated from the reference's own functions (oracle/make_golden.py). Only usable where
/root/reference exists (the authoring container); nothing on the GPU box imports this file.

The reference does not import as shipped. Neutralisations applied here, none of which
touch the arithmetic of the hot path:
  * sys.modules stubs for the absent packages `torch_geometric(.nn)` and `bottleneck`;
    `torch_geometric.nn.GCNConv` is RESTATED from the published PyG 2.5.3 algorithm (requirements.txt:53 pins
    torch_geometric==2.5.3): lin = Linear(in, out, bias=False) glorot-initialised, bias = zeros, gcn_norm with
    added self loops (weight 1), in-degree normalisation deg^-1/2[row] * w * deg^-1/2[col], aggregation at
    the target node (flow source_to_target), `+ bias`. No reference test pins this boundary -> the GCN part of
    the parity claim is "parity unpinned" (mitigated: only user rows are consumed, and those see only their
    own self loop, deg = 1, so no normalisation detail can change them — asserted in make_golden.py).
  * `Tensor.cuda` / `Module.cuda` become identity (gaussian_diffusion.py:744,889 and DNN.py:1157 call .cuda()
    unconditionally).
  * RNG recording: th.randint / th.randn_like / Tensor.multinomial / F.dropout are wrapped so the draws the
    reference consumed can be stored next to its outputs and re-injected into the restatement and the engine.
"""
from __future__ import annotations

import contextlib
import math
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


class _GCNConv(nn.Module):
    """Restatement of torch_geometric.nn.GCNConv (2.5.3) defaults: add_self_loops, normalize, bias."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        a = math.sqrt(6.0 / (in_channels + out_channels))  # glorot
        nn.init.uniform_(self.lin.weight, -a, a)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index):
        n = x.size(0)
        row, col = edge_index[0], edge_index[1]
        w = torch.ones(row.numel(), dtype=x.dtype, device=x.device)
        # add_remaining_self_loops: existing self loops keep their weight, every node ends with exactly one
        is_loop = row == col
        loop_w = torch.ones(n, dtype=x.dtype, device=x.device)
        loop_w[row[is_loop]] = w[is_loop]
        ar = torch.arange(n, device=x.device)
        row = torch.cat([row[~is_loop], ar])
        col = torch.cat([col[~is_loop], ar])
        w = torch.cat([w[~is_loop], loop_w])
        deg = torch.zeros(n, dtype=x.dtype, device=x.device).scatter_add_(0, col, w)
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        w = dis[row] * w * dis[col]
        h = self.lin(x)
        out = torch.zeros_like(h).index_add_(0, col, h[row] * w[:, None])
        return out + self.bias


def install_stubs():
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgnn = types.ModuleType("torch_geometric.nn")

        class _Empty(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        tgnn.LightGCN = _Empty
        tgnn.MessagePassing = _Empty
        tgnn.GCNConv = _GCNConv
        tg.nn = tgnn
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tgnn
    if "bottleneck" not in sys.modules:
        sys.modules["bottleneck"] = types.ModuleType("bottleneck")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self


def import_reference():
    """Returns (gaussian_diffusion, DNN, evaluate_utils, data_utils) modules of the reference."""
    install_stubs()
    import importlib
    gd = importlib.import_module("models.gaussian_diffusion")
    dnn = importlib.import_module("models.DNN")
    ev = importlib.import_module("evaluate_utils")
    du = importlib.import_module("data_utils")
    return gd, dnn, ev, du


def load_lightgcn_class(n_users: int, n_items: int):
    """lightGCN.py is a script (reads ml-100k at import, pdb at :249): extract only class LightGCN (:129-203)."""
    import ast
    import numpy as np
    import scipy.sparse as sp
    src = open(f"{REFERENCE_ROOT}/lightGCN.py").read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LightGCN"][0]
    mod = ast.Module(body=[cls], type_ignores=[])
    ns = {"nn": nn, "torch": torch, "sp": sp, "np": np, "n_users": n_users, "n_items": n_items}
    exec(compile(mod, "lightGCN.py", "exec"), ns)
    return ns["LightGCN"]


class Recorder:
    """Records every random draw the reference makes, in call order."""

    def __init__(self):
        self.randint, self.randn, self.multinomial, self.dropout = [], [], [], []


@contextlib.contextmanager
def record_rng(gd_module):
    rec = Recorder()
    th = gd_module.th
    o_randint, o_randn_like, o_multinomial, o_dropout = th.randint, th.randn_like, torch.Tensor.multinomial, F.dropout
    o_th_multinomial = th.multinomial

    def randint(*a, **k):
        r = o_randint(*a, **k)
        rec.randint.append(r.clone())
        return r

    def randn_like(x, *a, **k):
        r = o_randn_like(x, *a, **k)
        rec.randn.append(r.clone())
        return r

    def multinomial(self, *a, **k):
        r = o_multinomial(self, *a, **k)
        rec.multinomial.append(r.clone())
        return r

    def th_multinomial(x, *a, **k):
        r = o_th_multinomial(x, *a, **k)
        rec.multinomial.append(r.clone())
        return r

    def dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        keep = (torch.rand_like(x) >= p)
        rec.dropout.append(keep.clone())
        return x * keep.to(x.dtype) / (1.0 - p)

    th.randint, th.randn_like, torch.Tensor.multinomial, F.dropout = randint, randn_like, multinomial, dropout
    th.multinomial = th_multinomial
    try:
        yield rec
    finally:
        th.randint, th.randn_like, torch.Tensor.multinomial, F.dropout = o_randint, o_randn_like, o_multinomial, o_dropout
        th.multinomial = o_th_multinomial


def make_args(**kw):
    d = dict(user_guided=1, gcnLayerNum=2, noise_type=0)
    d.update(kw)
    return types.SimpleNamespace(**d)
