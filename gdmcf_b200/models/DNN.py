"""Denoisers of the GDMCF hot path — host-side mirror of the reference's models/DNN.py.

Same class names, constructor arguments, `forward` signatures and `state_dict` keys as the reference
(`DNN` models/DNN.py:11-88; `DNNOneHotEmbeddingGCN` :1105-1327 with `LayerGCN` :1077-1103 and
`nt_xent_loss` :479-508), so weights copy both ways with `load_state_dict`. The modules only own fp32
master parameters; every contraction runs in libgdmcf_sm100.so (tcgen05 GEMMs with fused epilogues)
on bf16 operand copies that are re-derived whenever a parameter changes. There is no torch compute path.

Differences by design:
  * the GCN is evaluated on the B user rows only: edges are user->item and GCNConv aggregates at the
    target, so user rows see only their self loop (weight 1) — item rows are dead work in the reference;
  * the last `emb_size` input columns of the first layers (`cat([x, emb])`) are folded into a per-timestep
    bias table [steps, d] instead of widening K to a non-multiple of 8;
  * at inference with `x_tU = one_hot(x0)` the one-hot encoder is a sparse gather-sum
    (`base + sum_{i in row} delta[i]`) and is hoisted out of the reverse loop.
`precision="fp32"` runs every contraction as a 3-segment hi/lo bf16 split (~1e-5); default "bf16" (~1e-3).
"""
from __future__ import annotations

import copy
import math
import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from ..kernels import Bf16Mat

_MAX_T_TABLE = 1024


def c_ok(hidden_channels: int) -> bool:
    return hidden_channels % 64 == 0


def _as_i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 else t.to(torch.int32)


class _OperandCache:
    """bf16 (hi[, lo]) operand copies and derived tables, keyed by name, invalidated by parameter version.
    `build(prev)` receives the stale value (or None) and refreshes it IN PLACE when shapes still match, so every
    operand keeps its address for the life of the model: a captured CUDA graph that contains the refresh kernels
    stays valid across optimizer steps."""

    def __init__(self):
        self._items: Dict[str, tuple] = {}
        self.epoch = 0  # bumped by optimizers that write parameters behind torch's version counter

    def get(self, name: str, params, build):
        ver = (self.epoch,) + tuple((p._version, p.data_ptr()) for p in params)
        hit = self._items.get(name)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = build(hit[1] if hit is not None else None)
        self._items[name] = (ver, val)
        return val

    def peek(self, name: str):
        """The cached value (possibly stale) or None."""
        hit = self._items.get(name)
        return hit[1] if hit is not None else None

    def mark_fresh(self, name: str, params) -> None:
        """The value was refreshed in place by someone else (FusedAdamW's fused refresh): adopt the current version."""
        hit = self._items.get(name)
        if hit is not None:
            self._items[name] = ((self.epoch,) + tuple((p._version, p.data_ptr()) for p in params), hit[1])

    def invalidate(self, prefix: str) -> None:
        """Mark the entries whose name starts with `prefix` stale; their storage is kept and refreshed in place."""
        for name, (ver, val) in list(self._items.items()):
            if name.startswith(prefix):
                self._items[name] = (None, val)

    def clear(self):
        self._items.clear()


class _EngineModule(nn.Module):
    """Shared plumbing: precision mode, operand cache, GEMM helper."""

    def _engine_init(self, precision: str):
        assert precision in ("bf16", "fp32")
        self.precision = precision
        self._ops = _OperandCache()
        self._bufs: Dict[tuple, object] = {}
        self._rng_calls = 0
        self.seed = 0

    @property
    def _lo(self) -> bool:
        return self.precision == "fp32"

    def weights_updated(self):
        """Called by optimizers that update parameters through raw pointers (gdmcf_b200.optim.FusedAdamW)."""
        self._ops.epoch += 1

    def _refresh_entry(self, param, op=None, op_t=None, inv=None, onehot=None, tcols=None, cols_used=0):
        """Description of the cached tensors derived from `param` for gdmcf_adamw_refresh: cache names + kernel arguments.
        Only tensors that have been built once (and therefore exist with the right shapes) take part."""
        names, kw = [], dict(cols_used=cols_used)
        rows, cols = param.shape
        for key, arg in ((op, "op"), (op_t, "op_t"), (inv, "inv_norm"), (tcols, "tcols")):
            val = self._ops.peek(key) if key else None
            if val is not None:
                kw[arg] = val
                names.append(key)
        tabs = self._ops.peek(onehot) if onehot else None
        if tabs is not None:
            kw["base"], kw["delta"] = tabs
            names.append(onehot)
        if "inv_norm" in kw or "delta" in kw:
            n = K.adamw_refresh_splits(rows, cols) * rows
            kw["rowpart"] = self._buf(("rowpart", id(param)), lambda: torch.empty(n, dtype=torch.float32, device=param.device))
        return (kw, names) if names else None

    def refresh_specs(self) -> Dict[int, tuple]:
        """id(param) -> (kwargs for kernels.adamw_refresh, cache names it refreshes). Overridden per backbone."""
        return {}

    def build_derived(self) -> None:
        """Build every tensor the contractions derive from the weights (training- AND inference-side). A captured training
        step refreshes, in its optimizer pass, exactly the derived tensors that exist when it is captured; an
        inference-only program captured later relies on that refresh. Overridden per backbone."""

    def invalidate_time_tables(self) -> None:
        """The per-timestep bias tables are derived from emb_layer and the first layers' time columns and are not part of
        the optimizer's fused refresh: rebuild them (in place) at the next use."""
        self._ops.invalidate("tb")

    def adopt_refreshed(self, param, names) -> None:
        for n in names:
            self._ops.mark_fresh(n, [param])

    def _next_offset(self) -> int:
        self._rng_calls += 1
        return (self._rng_calls & ((1 << 22) - 1)) << 40  # disjoint Philox counter ranges per call (bits 40-61)

    def _buf(self, key, make):
        b = self._bufs.get(key)
        if b is None:
            b = make()
            self._bufs[key] = b
        return b

    def _mm(self, a: Bf16Mat, b: Bf16Mat, m: int, n: int, k: int, **epi):
        """C[m,n] = a[m,k] @ b[n,k]^T (+ fused epilogue) in the module's precision mode (fp32 mode: hi/lo split segments;
        an operand without a lo part is exact in bf16, e.g. one-hot rows)."""
        if self._lo and a.lo is not None and b.lo is not None:
            K.gemm([a.hi, a.hi, a.lo], [b.hi, b.lo, b.hi], m, n, [k, k, k], **epi)
        elif self._lo and b.lo is not None:
            K.gemm([a.hi, a.hi], [b.hi, b.lo], m, n, [k, k], **epi)
        elif self._lo and a.lo is not None:
            K.gemm([a.hi, a.lo], [b.hi, b.hi], m, n, [k, k], **epi)
        else:
            K.gemm([a.hi], [b.hi], m, n, [k], **epi)

    def _weight_operand(self, name: str, param: torch.Tensor, cols: Optional[int] = None, transpose: bool = False) -> Bf16Mat:
        def build(prev):
            w = param.detach()
            if cols is not None:
                w = w[:, :cols]
            shape = (w.shape[1], w.shape[0]) if transpose else tuple(w.shape)
            if prev is not None and ((prev.rows, prev.cols) != shape or prev.hi.device != w.device or (prev.lo is None) == self._lo):
                prev = None
            return (K.cast_bf16_transpose if transpose else K.cast_bf16)(w, with_lo=self._lo, out=prev)
        return self._ops.get(name + (".T" if transpose else "") + self.precision, [param], build)

    def __getstate__(self):  # torch.save(model) (main.py:375): drop device caches
        st = self.__dict__.copy()
        st["_ops"], st["_bufs"] = _OperandCache(), {}
        return st


def _init_linear(layer: nn.Linear):
    """init_weights, models/DNN.py:39-70: N(0, sqrt(2/(fan_in+fan_out))) weights, N(0, 0.001) biases."""
    fan_out, fan_in = layer.weight.shape
    std = np.sqrt(2.0 / (fan_in + fan_out))
    layer.weight.data.normal_(0.0, std)
    layer.bias.data.normal_(0.0, 0.001)


class DNN(_EngineModule):
    """models/DNN.py:11-88 — `[x, emb(t)] -> tanh(Linear(n_item+e -> d)) -> Linear(d -> n_item)`; with more than one entry in
    dims: tanh after every in_layer and after every out_layer but the last (:79-86). The first in_layer and the last out_layer
    are the catalogue-wide contractions; the layers between them ("middle") are small."""

    def __init__(self, in_dims, out_dims, emb_size, time_type="cat", norm=False, dropout=0.5, precision="bf16"):
        super().__init__()
        self.in_dims, self.out_dims = in_dims, out_dims
        assert out_dims[0] == in_dims[-1], "In and out dimensions must equal to each other."
        if time_type != "cat":
            raise ValueError("Unimplemented timestep embedding type %s" % time_type)
        if len(in_dims) < 2 or len(out_dims) < 2:
            raise ValueError("in_dims / out_dims need at least one hidden width; got %s / %s" % (in_dims, out_dims))
        if any(w % 8 for w in list(in_dims[1:]) + list(out_dims[:-1])):
            raise NotImplementedError("every entry of dims must be a multiple of 8 (16-byte aligned operand rows)")
        if norm:
            raise NotImplementedError("norm=True (F.normalize on the input) is outside the configured hot path")
        self.time_type, self.time_emb_dim, self.norm = time_type, emb_size, norm
        self.emb_layer = nn.Linear(emb_size, emb_size)
        in_t = [in_dims[0] + emb_size] + list(in_dims[1:])
        self.in_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(in_t[:-1], in_t[1:])])
        self.out_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(out_dims[:-1], out_dims[1:])])
        self.drop = nn.Dropout(dropout)
        self.init_weights()
        self._engine_init(precision)

    def init_weights(self):
        for layer in list(self.in_layers) + list(self.out_layers) + [self.emb_layer]:
            _init_linear(layer)

    # -- operands ------------------------------------------------------------------------------
    @property
    def n_item(self) -> int:
        return self.in_dims[0]

    @property
    def hidden(self) -> int:
        """Output width of the first (catalogue-wide) layer."""
        return self.in_dims[1]

    @property
    def d_dec(self) -> int:
        """Input width of the last (catalogue-wide) out_layer; == hidden for dims=[d]."""
        return self.out_dims[-2]

    @property
    def deep(self) -> bool:
        return len(self.in_dims) > 2 or len(self.out_dims) > 2

    @property
    def _dec_name(self) -> str:
        return f"out_layers.{len(self.out_layers) - 1}"

    def _middle(self):
        """The small tanh layers between the two catalogue-wide contractions: in_layers[1:] then out_layers[:-1]."""
        return ([(f"in_layers.{i}", l) for i, l in enumerate(self.in_layers) if i > 0] +
                [(f"out_layers.{j}", l) for j, l in enumerate(self.out_layers) if j < len(self.out_layers) - 1])

    def _mid_forward(self, h: Bf16Mat, B: int, acts=None) -> Bf16Mat:
        """h <- tanh(layer(h)) through the middle layers; returns the decoder's input operand. acts (training): a list that
        receives (fp32, operand) of every layer output for the backward pass; inference reuses cached buffers."""
        cur, dev = h, h.hi.device
        for name, layer in self._middle():
            n_out, n_in = layer.weight.shape
            if acts is not None:
                f32, op = torch.empty(B, n_out, dtype=torch.float32, device=dev), Bf16Mat.empty(B, n_out, dev, self._lo)
                acts.append((f32, op))
            else:
                f32 = None
                op = self._buf(("mid", name, B), lambda: Bf16Mat.empty(B, n_out, dev, self._lo))
            self._mm(cur, self._weight_operand(name, layer.weight), B, n_out, n_in, act=K.ACT_TANH, bias=layer.bias.detach(),
                     out_f32=f32, out_bf16=op.hi, out_bf16_lo=op.lo)
            cur = op
        return cur

    def _tables(self, T: int):
        l0 = self.in_layers[0]
        return self._ops.get(f"tb{T}", [self.emb_layer.weight, self.emb_layer.bias, l0.weight, l0.bias],
                             lambda prev: K.time_bias_table(self.emb_layer.weight.detach(), self.emb_layer.bias.detach(),
                                                            l0.weight.detach(), self.n_item, l0.bias.detach(), T, out=prev))[0]

    def build_derived(self) -> None:
        self._weight_operand("in0", self.in_layers[0].weight, cols=self.n_item)
        self._weight_operand("out0", self.out_layers[-1].weight)

    def refresh_specs(self):
        pr, out = self.precision, {}
        w1, wo = self.in_layers[0].weight, self.out_layers[-1].weight
        for prm, ent in ((w1, self._refresh_entry(w1, op="in0" + pr, tcols="in_layers.0.tcols", cols_used=self.n_item)),
                         (wo, self._refresh_entry(wo, op="out0" + pr, op_t="out0.T" + pr))):
            if ent is not None:
                out[id(prm)] = ent
        return out

    # -- forward -------------------------------------------------------------------------------
    def _encode(self, x_op: Bf16Mat, B: int, ts, t_const: int, T: int, h_out: Bf16Mat, h_f32=None):
        w1 = self._weight_operand("in0", self.in_layers[0].weight, cols=self.n_item)
        self._mm(x_op, w1, B, self.hidden, self.n_item, act=K.ACT_TANH, bias=self._tables(T), ld_bias=self.hidden,
                 row_t=ts, t_const=t_const, out_bf16=h_out.hi, out_bf16_lo=h_out.lo, out_f32=h_f32)

    def _decode(self, h: Bf16Mat, B: int, out_f32, out_op: Optional[Bf16Mat] = None, **post):
        wo = self._weight_operand("out0", self.out_layers[-1].weight)
        self._mm(h, wo, B, self.n_item, self.d_dec, bias=self.out_layers[-1].bias.detach(), out_f32=out_f32,
                 out_bf16=out_op.hi if out_op is not None else None, out_bf16_lo=out_op.lo if out_op is not None else None,
                 **post)

    @torch.no_grad()
    def forward(self, x, timesteps):
        """x: fp32 [B, n_item] CUDA tensor; timesteps: int [B]. Returns fp32 [B, n_item] (models/DNN.py:72-88)."""
        K.require_cuda(x)
        B, I = x.shape
        assert I == self.n_item and timesteps.shape == (B,)
        ts = _as_i32(timesteps)
        x_op = self._buf(("x", B), lambda: Bf16Mat.empty(B, I, x.device, self._lo))
        p = self.drop.p if self.training else 0.0
        K.qsample_dropout(x, B, I, x_op, dropout_p=p, seed=self.seed, offset=self._next_offset())
        h = self._buf(("h", B), lambda: Bf16Mat.empty(B, self.hidden, x.device, self._lo))
        self._encode(x_op, B, ts, 0, _MAX_T_TABLE, h)
        out = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=x.device)
        self._decode(self._mid_forward(h, B), B, out)
        return out[:, :I]

    @torch.no_grad()
    def reverse_loop(self, x0_f32, B: int, index, steps_total: int, c1, c2, x0_op: Optional[Bf16Mat] = None,
                     csr=None, users=None, noise_hook=None):
        """The p_sample loop (models/gaussian_diffusion.py:695-752) for this backbone: x_t <- c1[t]*f(x_t,t) + c2[t]*x_t,
        t = T-1..0, with the posterior mean fused into the output GEMM's epilogue. x0_f32: [B, ld4] fp32 start
        (x_t = x_start, or its q_sample); returns the fp32 buffer [B, ld4] holding the final x_t."""
        I, dev = self.n_item, x0_f32.device
        ld4 = x0_f32.shape[1]
        bufs = self._buf(("rev", B), lambda: dict(
            xa=torch.empty(B, ld4, dtype=torch.float32, device=dev), xb=torch.empty(B, ld4, dtype=torch.float32, device=dev),
            xop=Bf16Mat.empty(B, I, dev, self._lo), h=Bf16Mat.empty(B, self.hidden, dev, self._lo)))
        x_op = x0_op
        if x_op is None:
            x_op = bufs["xop"]
            K.qsample_dropout(x0_f32, B, I, x_op)
        cur, nxt = x0_f32, bufs["xa"]
        for t in reversed(range(steps_total)):
            self._encode(x_op, B, None, t, steps_total, bufs["h"])
            last = t == 0
            self._decode(self._mid_forward(bufs["h"], B), B, nxt, None if last else bufs["xop"], c1=c1, c2=c2, xt=cur, t_const=t)
            if noise_hook is not None and not last:
                noise_hook(nxt, bufs["xop"], t)  # sampling_noise: x_{t-1} = mean + sigma[t] * z (gaussian_diffusion.py:745-750)
            x_op = bufs["xop"]
            cur, nxt = nxt, (bufs["xb"] if nxt is bufs["xa"] else bufs["xa"])
        return cur


class _GCNLin(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        a = math.sqrt(6.0 / (i + o))  # PyG Linear(weight_initializer='glorot')
        nn.init.uniform_(self.weight, -a, a)


class GCNConvParams(nn.Module):
    """Parameters of torch_geometric.nn.GCNConv with its state_dict keys (`bias`, `lin.weight`)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = _GCNLin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))


class LayerGCN(nn.Module):
    """models/DNN.py:1077-1103. Holds the parameters; DNNOneHotEmbeddingGCN evaluates it on the user rows (self loop
    only): gcnLayerNum == 2: W2 relu(W1 x + b1) + b2; == 1: W1 x + b1; == 0: not applied."""

    def __init__(self, in_channels, hidden_channels, out_channels, residual=False, args=None):
        super().__init__()
        if residual:
            raise NotImplementedError("residual=True is never used by the reference model")
        # models/DNN.py:1081-1085: gcnLayerNum == 1 -> one in -> out convolution; any other value builds conv1 / conv2
        # (gcnLayerNum == 0 builds them as well but the model never calls the GCN, :1278)
        self.n_layers = getattr(args, "gcnLayerNum", 2) if args is not None else 2
        if self.n_layers == 1:
            self.conv1 = GCNConvParams(in_channels, out_channels)
        else:
            self.conv1 = GCNConvParams(in_channels, hidden_channels)
            self.conv2 = GCNConvParams(hidden_channels, out_channels)


class DNNOneHotEmbeddingGCN(_EngineModule):
    """models/DNN.py:1105-1327 — the GDMCF denoiser (noise_type=0, gcnLayerNum=2)."""

    def __init__(self, in_dims, out_dims, emb_size, time_type="cat", norm=False, dropout=0.5, item_num=2810,
                 user_num=5949, args=None, precision="bf16"):
        super().__init__()
        self.args = args
        # ablation switches of the reference CLI (models/DNN.py:1236-1259, :1278)
        self.noise_type = getattr(args, "noise_type", 0) if args is not None else 0
        self.gcn_layers = getattr(args, "gcnLayerNum", 2) if args is not None else 2
        if self.noise_type not in (0, 1, 2):
            raise ValueError("noise_type must be 0, 1 or 2")
        in_dims, out_dims = list(in_dims), list(out_dims)
        self.in_dims = in_dims
        self.in_dims2 = copy.deepcopy(in_dims)
        self.in_dims2[0] *= 2
        self.out_dims = out_dims
        assert out_dims[0] == in_dims[-1], "In and out dimensions must equal to each other."
        if time_type != "cat":
            raise ValueError("Unimplemented timestep embedding type %s" % time_type)
        if len(in_dims) < 2:
            raise ValueError("in_dims needs the input width and at least one hidden width; got %s" % (in_dims,))
        if norm:
            raise NotImplementedError("norm=True is outside the configured hot path")
        if any(w % 8 for w in in_dims[1:]):
            raise NotImplementedError("every entry of dims must be a multiple of 8 (16-byte aligned operand rows / tower segments)")
        self.time_type, self.time_emb_dim, self.norm = time_type, emb_size, norm
        self.emb_layer = nn.Linear(emb_size, emb_size)
        in_t = [in_dims[0] + emb_size] + in_dims[1:]
        in_t2 = [self.in_dims2[0] + emb_size] + self.in_dims2[1:]
        out_t = out_dims
        out_t[0] += self.in_dims2[-1]  # models/DNN.py:1128 (mutates out_dims like the reference)
        self.in_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(in_t[:-1], in_t[1:])])
        self.in_layers2 = nn.ModuleList([nn.Linear(a, b) for a, b in zip(in_t2[:-1], in_t2[1:])])
        # out_layers is constructed by the reference but never used in forward (its grads stay None); it is kept
        # so that parameter counts and state_dict keys match.
        self.out_layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(out_t[:-1], out_t[1:])])
        self.drop = nn.Dropout(dropout)
        e_user = in_t[-1]
        e_item = in_t[-1] + e_user + in_t2[-1]
        self.embedding_item = nn.Embedding(item_num, e_item)
        self.embedding_user = nn.Embedding(user_num, e_user)
        self.gcn_model = LayerGCN(e_item, 512, e_item, residual=False, args=args)
        self.init_weights()
        self.sumW = nn.Parameter(torch.tensor(1.0))
        self.item_num, self.user_num = item_num, user_num
        # faithful graph mode: run LayerGCN over all B + I nodes with the edge set like the reference (item rows included)
        # instead of the user-row closed form; the results the model returns are the same
        self.faithful_graph = bool(getattr(args, "faithful_graph", False)) if args is not None else False
        self._edges = None
        self._engine_init(precision)

    def init_weights(self):
        for layer in list(self.in_layers) + list(self.in_layers2) + list(self.out_layers) + [self.emb_layer]:
            _init_linear(layer)
        nn.init.xavier_uniform_(self.embedding_item.weight)
        nn.init.xavier_uniform_(self.embedding_user.weight)

    # -- shapes --------------------------------------------------------------------------------
    @property
    def n_item(self) -> int:
        return self.in_dims[0]

    @property
    def hidden(self) -> int:
        """Width of h / h_U / e_user, the three segments of the user tower = the LAST entry of dims (models/DNN.py:1143-1147)."""
        return self.in_dims[-1]

    @property
    def d1(self) -> int:
        """Output width of the first (catalogue-wide) encoder layers; == hidden for dims=[d]."""
        return self.in_dims[1]

    @property
    def deep(self) -> bool:
        """dims has more than one entry: the encoders are tanh MLPs (models/DNN.py:1240-1252 loop over in_layers / in_layers2)."""
        return len(self.in_dims) > 2

    # -- operands ------------------------------------------------------------------------------
    def _tables(self, T: int):
        e = self.emb_layer
        l1, l2 = self.in_layers[0], self.in_layers2[0]
        tb1 = self._ops.get(f"tb1.{T}", [e.weight, e.bias, l1.weight, l1.bias],
                            lambda prev: K.time_bias_table(e.weight.detach(), e.bias.detach(), l1.weight.detach(), self.n_item,
                                                           l1.bias.detach(), T, out=prev))[0]
        tb2 = self._ops.get(f"tb2.{T}", [e.weight, e.bias, l2.weight, l2.bias],
                            lambda prev: K.time_bias_table(e.weight.detach(), e.bias.detach(), l2.weight.detach(),
                                                           2 * self.n_item, l2.bias.detach(), T, out=prev))[0]
        return tb1, tb2

    def _emb_table(self, T: int) -> torch.Tensor:
        """[T, e] time-embedding rows emb_layer(timestep_embedding(t)) — produced together with the bias tables."""
        self._tables(T)
        return self._ops.peek(f"tb1.{T}")[1]

    def _onehot_tables(self):
        w2 = self.in_layers2[0].weight
        def build(prev):
            d, I = self.d1, self.n_item
            ok = prev is not None and prev[0].shape == (d,) and prev[1].shape == (I, K.round_up(d, 4)) and prev[0].device == w2.device
            base = prev[0] if ok else torch.empty(d, dtype=torch.float32, device=w2.device)
            delta = prev[1] if ok else torch.zeros(I, K.round_up(d, 4), dtype=torch.float32, device=w2.device)
            # same producer as FusedAdamW's fused refresh: values agree bit for bit with the in-training tables
            K.refresh_derived(w2.detach(), cols_used=2 * I, delta=delta, base=base)
            return base, delta
        return self._ops.get("onehot", [w2], build)

    def _item_operands(self):
        E = self.embedding_item.weight
        e_op = self._weight_operand("E", E)
        def build_inv(prev):
            ok = prev is not None and prev.shape == (E.shape[0],) and prev.device == E.device
            inv = prev if ok else torch.empty(E.shape[0], dtype=torch.float32, device=E.device)
            K.refresh_derived(E.detach(), inv_norm=inv)  # same producer as FusedAdamW's fused refresh
            return inv
        inv = self._ops.get("E.inv", [E], build_inv)
        return e_op, inv

    def build_derived(self) -> None:
        self._weight_operand("in0", self.in_layers[0].weight, cols=self.n_item)
        self._weight_operand("in2", self.in_layers2[0].weight, cols=2 * self.n_item)
        if self.gcn_layers > 0:
            self._weight_operand("gcn1", self.gcn_model.conv1.lin.weight)
        if self.gcn_layers == 2:
            self._weight_operand("gcn2", self.gcn_model.conv2.lin.weight)
        self._onehot_tables()   # inference-side: sparse one-hot encoder tables
        self._item_operands()   # item table operand + inverse norms
        self._weight_operand("E", self.embedding_item.weight, transpose=True)

    def refresh_specs(self):
        pr, out = self.precision, {}
        w1, w2, E = self.in_layers[0].weight, self.in_layers2[0].weight, self.embedding_item.weight
        ents = [(w1, self._refresh_entry(w1, op="in0" + pr, tcols="in_layers.0.tcols", cols_used=self.n_item)),
                (w2, self._refresh_entry(w2, op="in2" + pr, onehot="onehot", tcols="in_layers2.0.tcols", cols_used=2 * self.n_item)),
                (E, self._refresh_entry(E, op="E" + pr, op_t="E.T" + pr, inv="E.inv"))]
        if self.gcn_layers > 0:
            g1 = self.gcn_model.conv1.lin.weight
            ents.append((g1, self._refresh_entry(g1, op="gcn1" + pr, op_t="gcn1.T" + pr)))
        if self.gcn_layers == 2:
            g2 = self.gcn_model.conv2.lin.weight
            ents.append((g2, self._refresh_entry(g2, op="gcn2" + pr, op_t="gcn2.T" + pr)))
        for prm, ent in ents:
            if ent is not None:
                out[id(prm)] = ent
        return out

    # -- pieces of the forward -----------------------------------------------------------------
    def _hc_buffers(self, B: int, dev):
        d = self.hidden
        return self._buf(("hc", B), lambda: dict(
            hc_f32=torch.empty(B, 3 * d, dtype=torch.float32, device=dev), hc=Bf16Mat.empty(B, 3 * d, dev, self._lo),
            g1=Bf16Mat.empty(B, 512, dev, self._lo), g2=torch.empty(B, 3 * d, dtype=torch.float32, device=dev),
            hcp=Bf16Mat.empty(B, 3 * d, dev, self._lo), inv_u=torch.empty(B, dtype=torch.float32, device=dev),
            xop=Bf16Mat.empty(B, self.n_item, dev, self._lo), S=torch.empty(B, self.d1, dtype=torch.float32, device=dev)))

    def _seg(self, bufs, s: int):
        """fp32 / bf16 views of segment s (h | h_U | e_user) of the concatenated user tower."""
        d = self.hidden
        hc = bufs["hc"]
        return (bufs["hc_f32"][:, s * d:(s + 1) * d], hc.hi[:, s * d:], hc.lo[:, s * d:] if hc.lo is not None else None)

    # -- deep encoders (dims with more than one entry) ------------------------------------------------
    def _deep_layers(self, branch: int):
        """The layers after the first of in_layers (branch 0) / in_layers2 (branch 1), with their parameter name prefixes."""
        name = "in_layers" if branch == 0 else "in_layers2"
        return [(f"{name}.{i}", l) for i, l in enumerate(getattr(self, name)) if i > 0]

    def _first_out(self, bufs, branch: int, B: int):
        """Where the first layer of an encoder writes tanh(W1 [x, emb] + b): straight into its segment of the user tower for
        dims=[d]; into a [B, d1] activation buffer that _deep_forward continues from otherwise. Returns (f32, hi, lo)."""
        if not self.deep:
            return self._seg(bufs, branch)
        acts = bufs.setdefault("acts", {})
        if branch not in acts or acts[branch][0][1].rows != B:
            dev = bufs["hc_f32"].device
            widths = [l.weight.shape[0] for _, l in self._deep_layers(branch)][:-1]
            with_lo = bufs["hc"].lo is not None
            acts[branch] = [(torch.empty(B, w, dtype=torch.float32, device=dev), Bf16Mat.empty(B, w, dev, with_lo))
                            for w in [self.d1] + widths]
        f32, op = acts[branch][0]
        return f32, op.hi, op.lo

    def _deep_forward(self, bufs, branch: int, B: int) -> None:
        """h <- tanh(layer(h)) for the layers after the first (models/DNN.py:1240-1242, :1249-1251); the last one writes its
        segment of the user tower. Activations stay in bufs["acts"][branch] (the training backward reads them)."""
        if not self.deep:
            return
        acts = bufs["acts"][branch]
        layers = self._deep_layers(branch)
        for li, (name, layer) in enumerate(layers):
            n_out, n_in = layer.weight.shape
            a_op = acts[li][1]
            if li == len(layers) - 1:
                f32, hi, lo = self._seg(bufs, branch)
            else:
                f32, hi, lo = acts[li + 1][0], acts[li + 1][1].hi, acts[li + 1][1].lo
            w = self._weight_operand(name, layer.weight)
            a_in = a_op if self._lo else Bf16Mat(a_op.hi, None, B, n_in)
            self._mm(a_in, w, B, n_out, n_in, act=K.ACT_TANH, bias=layer.bias.detach(), out_f32=f32, out_bf16=hi,
                     out_bf16_lo=lo)

    def _encode_x(self, bufs, x_op: Bf16Mat, B: int, ts, t_const: int, T: int):
        tb1, _ = self._tables(T)
        w1 = self._weight_operand("in0", self.in_layers[0].weight, cols=self.n_item)
        f32, hi, lo = self._first_out(bufs, 0, B)
        self._mm(x_op, w1, B, self.d1, self.n_item, act=K.ACT_TANH, bias=tb1, ld_bias=self.d1, row_t=ts,
                 t_const=t_const, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
        self._deep_forward(bufs, 0, B)

    def _encode_onehot_from_S(self, bufs, B: int, ts, t_const: int, T: int):
        _, tb2 = self._tables(T)
        f32, hi, lo = self._first_out(bufs, 1, B)
        K.bias_act_rows(bufs["S"], B, self.d1, bias=tb2, ld_bias=self.d1, row_t=ts, t_const=t_const,
                        act=K.ACT_TANH, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
        self._deep_forward(bufs, 1, B)

    def _x_branch_operand(self, x_op: Bf16Mat, xu_op_hi: Optional[torch.Tensor], B: int) -> Bf16Mat:
        """Input of the continuous encoder: x (models/DNN.py:1238), or with noise_type 1 the first n_item columns of the
        INTERLEAVED one-hot matrix [i0c0, i0c1, i1c0, ...] (:1236-1237) — a column slice of the one-hot operand."""
        if self.noise_type != 1:
            return x_op
        return Bf16Mat(xu_op_hi[:, : self.n_item], None, B, self.n_item)

    def _xx_operand(self, x_op: Bf16Mat, B: int) -> Bf16Mat:
        """noise_type 2: the one-hot encoder is fed cat([x, x]) (models/DNN.py:1246-1247)."""
        I = self.n_item
        xx = self._buf(("xx", B), lambda: Bf16Mat.empty(B, 2 * I, x_op.hi.device, self._lo))
        for src, dst in ((x_op.hi, xx.hi), (x_op.lo, xx.lo)):
            if src is not None:
                dst[:, :I].copy_(src[:, :I])
                dst[:, I:2 * I].copy_(src[:, :I])
        return xx

    def _encode_onehot_dense(self, bufs, xu_op_hi: torch.Tensor, B: int, ts, t_const: int, T: int, to_S: bool, lo=None):
        """h_U from a dense one-hot-branch operand [B, 2I] (training / noised inference): tensor-core GEMM. lo: residual of
        the operand in fp32 mode when it is not exact in bf16 (noise_type 2 feeds continuous values)."""
        w2 = self._weight_operand("in2", self.in_layers2[0].weight, cols=2 * self.n_item)
        k, d1 = 2 * self.n_item, self.d1
        if lo is not None and self._lo:
            assert not to_S
            _, tb2 = self._tables(T)
            f32, hi, lo_out = self._first_out(bufs, 1, B)
            K.gemm([xu_op_hi, xu_op_hi, lo], [w2.hi, w2.lo, w2.hi], B, d1, [k, k, k], act=K.ACT_TANH, bias=tb2,
                   ld_bias=d1, row_t=ts, t_const=t_const, out_f32=f32, out_bf16=hi, out_bf16_lo=lo_out)
            self._deep_forward(bufs, 1, B)
            return
        if to_S:  # pre-activation only (step-invariant part, hoisted out of the reverse loop)
            K.gemm([xu_op_hi], [w2.hi], B, d1, [k], out_f32=bufs["S"]) if not self._lo else \
                K.gemm([xu_op_hi, xu_op_hi], [w2.hi, w2.lo], B, d1, [k, k], out_f32=bufs["S"])
            return
        _, tb2 = self._tables(T)
        f32, hi, lo = self._first_out(bufs, 1, B)
        epi = dict(act=K.ACT_TANH, bias=tb2, ld_bias=d1, row_t=ts, t_const=t_const, out_f32=f32, out_bf16=hi,
                   out_bf16_lo=lo)
        if self._lo:  # the one-hot operand is exact in bf16 ({0,1,2}): only the weight needs the lo term
            K.gemm([xu_op_hi, xu_op_hi], [w2.hi, w2.lo], B, d1, [k, k], **epi)
        else:
            K.gemm([xu_op_hi], [w2.hi], B, d1, [k], **epi)
        self._deep_forward(bufs, 1, B)

    def _user_tower(self, bufs, B: int):
        """GCN on the user rows + sumW mix + row norms (models/DNN.py:1274-1288, :1320)."""
        d3 = 3 * self.hidden
        edges = getattr(self, "_edges", None)
        if self.faithful_graph and edges is not None and self.gcn_layers > 0:
            g_all = self.gcn_all_nodes(bufs["hc"], B, edges)
            K.mix_rownorm(bufs["hc_f32"], B, d3, g=g_all[:B], sumw=self.sumW.detach(), out=bufs["hcp"], inv_norm=bufs["inv_u"])
            return
        if self.gcn_layers == 0:  # no GCN (:1278): hc * sumW + hc * (1 - sumW)
            K.mix_rownorm(bufs["hc_f32"], B, d3, g=bufs["hc_f32"], sumw=self.sumW.detach(), out=bufs["hcp"], inv_norm=bufs["inv_u"])
            return
        c1 = self.gcn_model.conv1
        wc1 = self._weight_operand("gcn1", c1.lin.weight)
        if self.gcn_layers == 1:  # one convolution 3d -> 3d, no activation (:1095-1096)
            self._mm(bufs["hc"], wc1, B, d3, d3, bias=c1.bias.detach(), out_f32=bufs["g2"])
            K.mix_rownorm(bufs["hc_f32"], B, d3, g=bufs["g2"], sumw=self.sumW.detach(), out=bufs["hcp"], inv_norm=bufs["inv_u"])
            return
        c2 = self.gcn_model.conv2
        wc2 = self._weight_operand("gcn2", c2.lin.weight)
        if self._fused_tower():
            # both linears, the mix and the norms in one launch (csrc/tower.cu)
            K.user_tower(bufs["hc"], bufs["hc_f32"], wc1, c1.bias.detach(), wc2, c2.bias.detach(), self.sumW.detach(), B,
                         out=bufs["hcp"], inv_u=bufs["inv_u"])
            return
        g1 = bufs["g1"]
        self._mm(bufs["hc"], wc1, B, 512, d3, act=K.ACT_RELU, bias=c1.bias.detach(), out_bf16=g1.hi, out_bf16_lo=g1.lo)
        self._mm(g1, wc2, B, d3, 512, bias=c2.bias.detach(), out_f32=bufs["g2"])
        K.mix_rownorm(bufs["hc_f32"], B, d3, g=bufs["g2"], sumw=self.sumW.detach(), out=bufs["hcp"], inv_norm=bufs["inv_u"])

    # -- projected reverse loop ------------------------------------------------------------------------
    def _projection_operand(self) -> Bf16Mat:
        """P = W1[:, :I] diag(1 / ||E_i||) E   [d, 3d], K-major bf16 operand (hi[, lo]).
        The reverse step x_{t-1} = c1[t] * s_t + c2[t] * x_t (models/gaussian_diffusion.py:1047-1050) only ever re-enters
        the model through the LINEAR first layer h = tanh(W1 [x, emb] + b) (models/DNN.py:1240-1242), and the scores are
        s_t = ru * (hc' E^T) * ri (:1320-1325). Pushing the recurrence through W1:
            W1 x_{t-1} = c1[t] * ru * (hc' P^T) + c2[t] * (W1 x_t)
        so the intermediate steps need a [B, 3d] x [3d, d] product instead of the catalogue-wide scorer [B, 3d] x [3d, I]
        AND encoder [B, I] x [I, d]; only the last step (c1[0] = 1, c2[0] = 0: its output IS x_0) runs the full scorer.
        Same arithmetic as the reference up to rounding order; P is rebuilt whenever W1 or E change (one 2*d*3d*I
        contraction per weight version)."""
        W1, E = self.in_layers[0].weight, self.embedding_item.weight
        d, d3, I = self.d1, 3 * self.hidden, self.n_item

        def build(prev):
            _, inv_i = self._item_operands()
            eT = self._weight_operand("E", E, transpose=True)                      # [3d, I]
            w1r = self._buf(("w1r", self.precision), lambda: Bf16Mat.empty(d, I, W1.device, self._lo))
            K.scale_cols_cast(W1.detach(), I, inv_i, with_lo=self._lo, out=w1r)      # W1 diag(ri)
            ok = prev is not None and (prev.rows, prev.cols) == (d, d3) and prev.hi.device == W1.device and (prev.lo is None) != self._lo
            P = prev if ok else Bf16Mat.empty(d, d3, W1.device, self._lo)
            self._mm(w1r, eT, d, d3, I, out_bf16=P.hi, out_bf16_lo=P.lo)
            return P
        return self._ops.get("proj" + self.precision, [W1, E], build)

    def refresh_inference_operands(self) -> None:
        """Rebuild (in place) the inference-only operands that no optimizer pass refreshes: the projection P. Called by
        engine.StepEngine.flush() before inference-only captured programs run on weights that training has changed."""
        if self.can_project() and self._ops.peek("proj" + self.precision) is not None:
            self._ops.invalidate("proj")
            with torch.no_grad():
                self._projection_operand()

    def can_project(self) -> bool:
        """The projected reverse loop covers the configured model (noise_type 0, closed-form user rows)."""
        return self.noise_type == 0 and not self.faithful_graph and os.environ.get("GDMCF_PROJECTED_LOOP", "1") != "0"

    @torch.no_grad()
    def reverse_loop_projected(self, x0_f32, B: int, index, steps_total: int, c1, c2, x0_op: Optional[Bf16Mat] = None,
                               csr=None, users=None, xu_op: Optional[torch.Tensor] = None):
        """The p_sample loop (models/gaussian_diffusion.py:695-752) with the recurrence carried in the encoder's
        pre-activation space (see _projection_operand). Requires c1[0] = 1, c2[0] = 0 (START_X, every schedule: the last
        reverse step returns the model output) and no per-step noise. Returns the fp32 buffer [B, ld4] holding x_0."""
        I, d, dh, dev = self.n_item, self.d1, self.hidden, x0_f32.device
        ld4 = x0_f32.shape[1]
        bufs = self._hc_buffers(B, dev)
        rb = self._buf(("proj", B, ld4), lambda: dict(out=torch.empty(B, ld4, dtype=torch.float32, device=dev),
                                                       pa=torch.empty(B, d, dtype=torch.float32, device=dev),
                                                       pb=torch.empty(B, d, dtype=torch.float32, device=dev)))
        if xu_op is not None:
            self._encode_onehot_dense(bufs, xu_op, B, None, 0, steps_total, to_S=True)
        else:
            base, delta = self._onehot_tables()
            K.encode_onehot_gather(csr[0], csr[1], users, B, base, delta, d, bufs["S"])
        if getattr(self, "_user_rows_staged", False):
            self._user_rows_staged = False
        else:
            f32, hi, lo = self._seg(bufs, 2)
            K.gather_rows(self.embedding_user.weight.detach(), index, B, dh, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
        x_op = x0_op
        if x_op is None:
            x_op = bufs["xop"]
            K.qsample_dropout(x0_f32, B, I, x_op)
        # W1 x_{T-1}: the only catalogue-wide encoder product of the loop
        w1 = self._weight_operand("in0", self.in_layers[0].weight, cols=I)
        cur, nxt = rb["pa"], rb["pb"]
        self._mm(x_op, w1, B, d, I, out_f32=cur)
        P = self._projection_operand()
        tb1, _ = self._tables(steps_total)
        for t in reversed(range(steps_total)):
            f32, hi, lo = self._first_out(bufs, 0, B)
            K.bias_act_rows(cur, B, d, bias=tb1, ld_bias=d, t_const=t, act=K.ACT_TANH, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
            self._deep_forward(bufs, 0, B)
            self._encode_onehot_from_S(bufs, B, None, t, steps_total)
            self._user_tower(bufs, B)
            if t == 0:
                self._score(bufs, B, rb["out"])          # x_0 = the model output of the last step
            else:
                self._mm(bufs["hcp"], P, B, d, 3 * dh, row_scale=bufs["inv_u"], c1=c1, c2=c2, xt=cur, t_const=t, out_f32=nxt)
                cur, nxt = nxt, cur
        return rb["out"]

    # -- faithful graph mode --------------------------------------------------
    @torch.no_grad()
    def gcn_all_nodes(self, hc: Bf16Mat, B: int, edges: torch.Tensor) -> torch.Tensor:
        """LayerGCN over ALL B + I nodes with the user -> item edge set, as the reference runs it (models/DNN.py:1217-1219,
        1277-1280, 1093-1103): GCNConv = self loops + symmetric in-degree normalisation + aggregation at the target node,
        here as tcgen05 contractions over all rows + the CSR SpMM kernel (gdmcf_spmm_csr_f32) for the aggregation.
        edges: [B, I] (non-zero = edge user b -> item i). Returns fp32 [B + I, 3d]; rows [:B] are the user rows the model
        consumes (identical to the self-loop-only closed form of the default path), rows [B:] the item rows it discards."""
        I, d3, dev = self.n_item, 3 * self.hidden, hc.hi.device
        N = B + I
        G = edges.bool()
        cnt = G.sum(0)                                             # in-degree of every item node (without the self loop)
        deg_item = (cnt + 1).float()
        pairs = G.t().nonzero()                                    # [E, 2] = (item, user), sorted by item then user
        E_n = pairs.shape[0]
        rowptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        rowptr[1:B + 1] = torch.arange(1, B + 1, device=dev)       # user rows: the self loop only
        rowptr[B + 1:] = B + torch.cumsum(cnt + 1, 0)
        col = torch.empty(N + E_n, dtype=torch.int32, device=dev)
        val = torch.empty(N + E_n, dtype=torch.float32, device=dev)
        col[:B] = torch.arange(B, dtype=torch.int32, device=dev)
        val[:B] = 1.0                                              # deg = 1: weight 1
        starts = rowptr[B:B + I]
        col[starts] = (B + torch.arange(I, device=dev)).int()
        val[starts] = 1.0 / deg_item                               # self loop of an item: deg^-1/2 * deg^-1/2
        if E_n:
            item_of = pairs[:, 0]
            first = torch.cumsum(cnt, 0) - cnt
            pos = starts[item_of] + 1 + (torch.arange(E_n, device=dev) - first[item_of])
            col[pos] = pairs[:, 1].int()
            val[pos] = deg_item[item_of].rsqrt()                   # user (deg 1) -> item: 1 * deg_item^-1/2
        plan = K.spmm_plan(rowptr.to(torch.int32).cpu(), chunk=128, device=dev)
        e_op, _ = self._item_operands()
        c1 = self.gcn_model.conv1
        wc1 = self._weight_operand("gcn1", c1.lin.weight)
        w_out = c1.lin.weight.shape[0]                             # hidden (2 layers) or 3d (1 layer)
        wp = K.round_up(w_out, 64)                                 # the SpMM kernel walks 64-column slabs
        X1 = torch.zeros(N, wp, dtype=torch.float32, device=dev)
        self._mm(hc, wc1, B, w_out, d3, out_f32=X1[:B])            # x W^T on the user rows ...
        self._mm(e_op, wc1, I, w_out, d3, out_f32=X1[B:])          # ... and on the item rows (the reference's dead 211 GF)
        Y1 = K.spmm_csr(plan, col, val, X1)                        # aggregation at the target nodes
        bias1 = torch.zeros(wp, dtype=torch.float32, device=dev)
        bias1[:w_out] = c1.bias.detach()
        if self.gcn_layers == 1:
            out = Y1 + bias1
            self.last_gcn_all = out[:, :w_out]
            return self.last_gcn_all
        R = torch.empty_like(Y1)
        K.bias_act_rows(Y1, N, wp, bias=bias1, act=K.ACT_RELU, out_f32=R)   # relu, then LeakyReLU(0.1) = identity (:1097-1098)
        Z = K.spmm_csr(plan, col, val, R)                          # A'(R W2^T) = (A' R) W2^T: aggregate in the narrow space
        c2 = self.gcn_model.conv2
        wc2 = self._weight_operand("gcn2", c2.lin.weight)
        Zop = K.cast_bf16(Z[:, :w_out], with_lo=self._lo)
        out = torch.empty(N, K.round_up(d3, 4), dtype=torch.float32, device=dev)
        self._mm(Zop, wc2, N, d3, w_out, bias=c2.bias.detach(), out_f32=out)
        self.last_gcn_all = out[:, :d3]
        return self.last_gcn_all

    def _fused_tower(self) -> bool:
        """bf16 mode uses the one-launch tower kernel (GDMCF_FUSED_TOWER=0: the two contractions + mix kernels)."""
        return (not self._lo and self.gcn_layers == 2 and os.environ.get("GDMCF_FUSED_TOWER", "1") != "0"
                and c_ok(self.gcn_model.conv1.lin.weight.shape[0]))

    @property
    def needs_dense_onehot(self) -> bool:
        """noise_type 1 feeds columns of the interleaved one-hot matrix to the continuous encoder: the dense operand must
        exist even when x_tU = one_hot(x0) (no sparse shortcut)."""
        return self.noise_type == 1

    def _score(self, bufs, B: int, out_f32, out_op: Optional[Bf16Mat] = None, **post):
        """cosine_similarity_cuda (models/DNN.py:1304-1327) with the norms applied in the GEMM epilogue."""
        e_op, inv_i = self._item_operands()
        self._mm(bufs["hcp"], e_op, B, self.n_item, 3 * self.hidden, row_scale=bufs["inv_u"], col_scale=inv_i,
                 out_f32=out_f32, out_bf16=out_op.hi if out_op is not None else None,
                 out_bf16_lo=out_op.lo if out_op is not None else None, **post)

    @torch.no_grad()
    def stage_user_rows(self, B: int, index) -> None:
        """Gather embedding_user[index] into the user-tower buffer ahead of the next reverse_loop call (which then skips
        its own gather). engine.StepEngine uses it to update the user table on another stream while the loop runs."""
        bufs = self._hc_buffers(B, index.device)
        f32, hi, lo = self._seg(bufs, 2)
        K.gather_rows(self.embedding_user.weight.detach(), _as_i32(index), B, self.hidden, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
        self._user_rows_staged = True

    # -- reference call surface ----------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, timesteps, x_U, index=None, graph=None, RCloss=False):
        """models/DNN.py:1207-1302. x: fp32 [B, I]; x_U: [B, I, 2]; index: user ids [B]. `graph` only feeds item rows of
        the GCN, which the model never reads (see module docstring): it is ignored unless `faithful_graph` is set.
        Inference/eval forward (no autograd); training goes through GaussianDiffusionDiscrete.training_losses."""
        if RCloss:
            raise NotImplementedError("RCloss is produced by the fused training step (training_losses), not by forward()")
        K.require_cuda(x, x_U)
        B, I = x.shape
        assert I == self.n_item and timesteps.shape == (B,) and index is not None
        dev = x.device
        # ct = graph.argmax(2); edges = nonzero(ct) (models/DNN.py:1217-1219) — only consumed in faithful graph mode
        self._edges = graph.argmax(dim=2).to(torch.uint8) if (self.faithful_graph and graph is not None) else None
        ts = _as_i32(timesteps)
        bufs = self._hc_buffers(B, dev)
        p = self.drop.p if self.training else 0.0
        K.qsample_dropout(x, B, I, bufs["xop"], dropout_p=p, seed=self.seed, offset=self._next_offset())
        xu = self._buf(("xu", B), lambda: torch.zeros(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device=dev))
        xu_f = x_U.reshape(B, 2 * I).float()
        tmp = Bf16Mat(xu, None, B, 2 * I)
        K.qsample_dropout(xu_f, B, 2 * I, tmp, dropout_p=p, seed=self.seed + 1, offset=self._next_offset())
        self._encode_x(bufs, self._x_branch_operand(bufs["xop"], xu, B), B, ts, 0, _MAX_T_TABLE)
        if self.noise_type == 2:
            xx = self._xx_operand(bufs["xop"], B)
            self._encode_onehot_dense(bufs, xx.hi, B, ts, 0, _MAX_T_TABLE, to_S=False, lo=xx.lo)
        else:
            self._encode_onehot_dense(bufs, xu, B, ts, 0, _MAX_T_TABLE, to_S=False)
        f32, hi, lo = self._seg(bufs, 2)
        K.gather_rows(self.embedding_user.weight.detach(), _as_i32(index.to(dev)), B, self.hidden, out_f32=f32, out_bf16=hi,
                      out_bf16_lo=lo)
        self._user_tower(bufs, B)
        out = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
        self._score(bufs, B, out)
        return out[:, :I]

    @torch.no_grad()
    def reverse_loop(self, x0_f32, B: int, index, steps_total: int, c1, c2, x0_op: Optional[Bf16Mat] = None,
                     csr=None, users=None, xu_op: Optional[torch.Tensor] = None, noise_hook=None, graph_hook=None):
        """The p_sample loop (models/gaussian_diffusion.py:695-752) for GDMCF. The one-hot encoder's pre-activation
        S(x_tU) does not depend on t and is computed once: sparse gather from the CSR rows when x_tU = one_hot(x0)
        (csr=(rowptr, col), users), else one dense GEMM on `xu_op` [B, 2I]. Per step: encoder GEMM (split-K) ->
        user tower -> scorer GEMM whose epilogue applies the cosine norms and the posterior mean."""
        I, d, dev = self.n_item, self.hidden, x0_f32.device
        ld4 = x0_f32.shape[1]
        bufs = self._hc_buffers(B, dev)
        rb = self._buf(("rev", B, ld4), lambda: dict(xa=torch.empty(B, ld4, dtype=torch.float32, device=dev),
                                                     xb=torch.empty(B, ld4, dtype=torch.float32, device=dev)))
        nt = self.noise_type
        if nt == 2:
            pass  # the one-hot encoder is fed [x_t, x_t]: nothing step-invariant to hoist
        elif xu_op is not None:
            self._encode_onehot_dense(bufs, xu_op, B, None, 0, steps_total, to_S=True)
        else:
            base, delta = self._onehot_tables()
            K.encode_onehot_gather(csr[0], csr[1], users, B, base, delta, self.d1, bufs["S"])
        if getattr(self, "_user_rows_staged", False):
            self._user_rows_staged = False  # gathered ahead of the loop by stage_user_rows()
        else:
            f32, hi, lo = self._seg(bufs, 2)
            K.gather_rows(self.embedding_user.weight.detach(), index, B, d, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
        x_op = x0_op
        if x_op is None:
            x_op = bufs["xop"]
            K.qsample_dropout(x0_f32, B, I, x_op)
        cur, nxt = x0_f32, rb["xa"]
        for t in reversed(range(steps_total)):
            self._encode_x(bufs, self._x_branch_operand(x_op, xu_op, B), B, None, t, steps_total)
            if nt == 2:
                xx = self._xx_operand(x_op, B)
                self._encode_onehot_dense(bufs, xx.hi, B, None, t, steps_total, to_S=False, lo=xx.lo)
            else:
                self._encode_onehot_from_S(bufs, B, None, t, steps_total)
            # faithful graph mode: the diffusion's random edge bookkeeping of this step (gaussian_diffusion.py:710-729)
            self._edges = graph_hook(t) if (graph_hook is not None and self.faithful_graph) else None
            self._user_tower(bufs, B)
            last = t == 0
            self._score(bufs, B, nxt, None if last else bufs["xop"], c1=c1, c2=c2, xt=cur, t_const=t)
            if noise_hook is not None and not last:
                noise_hook(nxt, bufs["xop"], t)  # sampling_noise (gaussian_diffusion.py:745-750)
            x_op = bufs["xop"]
            cur, nxt = nxt, (rb["xb"] if nxt is rb["xa"] else rb["xa"])
        return cur


def timestep_embedding(timesteps, dim, max_period=10000):
    """models/DNN.py:1806-1825 (host-side helper kept for API parity; the kernels build their own tables)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        embedding = torch.cat([embedding, torch.zeros_like(embedding[:, :1])], dim=-1)
    return embedding
