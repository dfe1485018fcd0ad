"""Fused training step behind GaussianDiffusionDiscrete.training_losses (models/gaussian_diffusion.py:834-957).

The reference builds the loss with autograd over ATen ops. Here one torch.autograd.Function wraps the whole
denoiser: its forward launches the CUDA forward (noising, both encoders, nt_xent logits, user tower, cosine
scorer, per-row MSE) and returns `(mse[B], closs)`; its backward launches the hand-written backward (loss-gradient
pass, dgrad/wgrad contractions on the tensor cores, activation/mix/softmax backward) and hands one gradient per
parameter back to autograd. The O(B) bookkeeping of training_losses (SNR reweighting, Lt_history, /pt, +0.1*closs)
stays in torch on [B]-sized vectors so dtypes follow the reference (float64 loss). The caller keeps its loop:
    losses = diffusion.training_losses(model, batch, reweight, index=index); losses["loss"].mean().backward()
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import kernels as K
from .kernels import Bf16Mat
from .models.DNN import DNN, DNNOneHotEmbeddingGCN, timestep_embedding


def _vec_cache(model, name, make):
    return model._buf(("vec", name), make)


def _ones(model, n, dev):
    return _vec_cache(model, f"ones{n}", lambda: torch.ones(n, dtype=torch.float32, device=dev))


def _arange32(model, n, dev):
    return _vec_cache(model, f"arange{n}", lambda: torch.arange(n, dtype=torch.int32, device=dev))


def _temb_table(model, T, dev):
    return _vec_cache(model, f"temb{T}", lambda: timestep_embedding(torch.arange(T), model.time_emb_dim).to(dev).contiguous())


def _time_cols(model, name, W, n_in):
    """Contiguous [d, e] copy of a first layer's time-embedding columns (cached per weight version): the strided
    view W[:, n_in:] would make every lane of the skinny contraction touch a different 137 KB-apart row."""
    def build(prev):
        src = W.detach()[:, n_in:]
        if prev is not None and prev.shape == src.shape and prev.device == src.device:
            prev.copy_(src)
            return prev
        return src.contiguous()
    return model._ops.get(name + ".tcols", [W], build)


def _grad_buffer(W: torch.Tensor, model=None, name: str = "") -> torch.Tensor:
    """fp32 gradient buffer for a [d, n] weight whose rows start 16 B aligned, so that the wgrad contraction can use the
    TMA-store epilogue (nn.Linear(n_item + emb_size, d) weights have an odd row length). Returned as a [d, n] view;
    autograd compacts it when it installs .grad. A caller-owned buffer registered in model._grad_views[name] (the
    data-parallel engine's persistent, row-padded reduce-scatter buffers) takes precedence."""
    views = getattr(model, "_grad_views", None) if model is not None else None
    if views and name in views:
        return views[name]
    d, n = W.shape
    ld = K.round_up(n, 4)
    if ld == n:
        return torch.empty_like(W)
    return torch.empty(d, ld, dtype=W.dtype, device=W.device)[:, :n]


def _mm_auto(model, a: Bf16Mat, b: Bf16Mat, m, n, k, **epi):
    """Contraction in the model's precision; operands that carry no lo part fall back to their hi part only."""
    if model._lo and a.lo is not None and b.lo is not None:
        K.gemm([a.hi, a.hi, a.lo], [b.hi, b.lo, b.hi], m, n, [k, k, k], **epi)
    elif model._lo and b.lo is not None:
        K.gemm([a.hi, a.hi], [b.hi, b.lo], m, n, [k, k], **epi)
    elif model._lo and a.lo is not None:
        K.gemm([a.hi, a.lo], [b.hi, b.hi], m, n, [k, k], **epi)
    else:
        K.gemm([a.hi], [b.hi], m, n, [k], **epi)


def _mm3(a: Bf16Mat, b: Bf16Mat, m, n, k, **epi):
    """Always split-precision (used for the nt_xent logits, whose softmax amplifies operand rounding 10x)."""
    K.gemm([a.hi, a.hi, a.lo], [b.hi, b.lo, b.hi], m, n, [k, k, k], **epi)


def _bf16_T(x: Bf16Mat, rows, cols, dev, extra_rows: int = 0) -> Bf16Mat:
    """Transpose a bf16 operand [rows, cols] -> [cols (+ extra_rows), rows] (hi and lo); the extra rows are left for
    the caller to fill (the time-embedding rows appended to a first layer's wgrad operand)."""
    out = Bf16Mat.empty(cols + extra_rows, rows, dev, x.lo is not None, zero=False)
    K.transpose_bf16(x.hi, rows, cols, out.hi)
    if x.lo is not None:
        K.transpose_bf16(x.lo, rows, cols, out.lo)
    return out


def _input_T_with_time_rows(model, A: Bf16Mat, emb_rows, B, n_in, e, dev) -> Bf16Mat:
    """[n_in + e, B] wgrad operand of a first layer: the transposed (noised, dropped-out) input rows followed by the
    transposed time-embedding rows emb(t_b) — `cat([x, emb])` of models/DNN.py:79/1240/1250 seen from the weight gradient.
    One contraction then yields all n_in + e gradient columns (the e time columns used to be a separate 23 us kernel)."""
    AT = _bf16_T(A, B, n_in, dev, extra_rows=e)
    tail = Bf16Mat(AT.hi[n_in:], AT.lo[n_in:] if AT.lo is not None else None, e, B)
    K.cast_bf16_transpose(emb_rows, with_lo=AT.lo is not None, out=tail)
    return AT


def _hand_over(model, grads: Dict[str, torch.Tensor], names):
    """Gradients for autograd, in parameter order. A padded wgrad buffer (see _grad_buffer) is installed as .grad directly:
    autograd's AccumulateGrad would compact it with a 137 MB copy; FusedAdamW and the all-reduce read it through its
    leading dimension."""
    params = dict(model.named_parameters())
    out = []
    for n in names:
        g = grads.get(n)  # None: the parameter takes no part in this configuration (gcnLayerNum 0: the GCN is never called)
        if g is None:
            out.append(None)
            continue
        if g.dim() == 2 and not g.is_contiguous():
            prm = params[n]
            if prm.grad is None:
                prm.grad = g
            else:
                prm.grad.add_(g)
            g = None
        out.append(g)
    return tuple(out)


class _Ctx:
    """Tensors the backward needs (kept alive between forward and backward of one step)."""


def _is_eps(diff) -> bool:
    return getattr(diff.mean_type, "name", "") == "EPSILON"


def _noised_input(diff, model, c, x0, B, I, ts, inject, p):
    """q_sample + dropout + operand cast (gaussian_diffusion.py:868-870, DNN.py:78/1232) -> c.A1, and the regression target
    of the loss: x_start (START_X) or the noise (EPSILON, :895-898). With EPSILON the rows drawn at t == 0 use the
    likelihood term mean((x0 - pred_xstart)^2 / 2) (:924-928), which is the same squared distance to the model output
    after a change of variables: (srm1^2 / 2) * mean((out - (sr * x_t - x0) / srm1)^2). c.loss_scale carries the factor."""
    dev = x0.device
    c.A1 = Bf16Mat.empty(B, I, dev, model._lo, zero=False)
    keep = inject.get("keep_x")
    noise = inject.get("noise")
    eps_mode = _is_eps(diff)
    xt = None
    if eps_mode:
        if noise is None:
            noise = torch.randn(B, I, device=dev)
        xt = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
    K.qsample_dropout(x0, B, I, c.A1, row_t=ts, sqrt_ab=diff._f32["sqrt_alphas_cumprod"],
                      sqrt_1mab=diff._f32["sqrt_one_minus_alphas_cumprod"],
                      noise=noise.float().contiguous() if noise is not None else None,
                      keep=keep.to(torch.uint8).contiguous() if keep is not None else None, dropout_p=p, seed=diff.seed,
                      offset=diff._offset(), epoch=diff._epoch, xt_out=xt)
    c.target, c.loss_scale = x0, None
    if eps_mode:
        sr0 = float(diff.sqrt_recip_alphas_cumprod[0])
        srm0 = float(diff.sqrt_recipm1_alphas_cumprod[0])
        t0 = (ts == 0)
        tgt = torch.zeros(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
        tgt[:, :I] = torch.where(t0[:, None], (sr0 * xt[:, :I] - x0[:, :I]) / srm0, noise.float())
        c.target = tgt
        c.loss_scale = torch.where(t0, torch.full((), srm0 * srm0 / 2.0, device=dev), torch.ones((), device=dev)).float()


def _row_mse(c, B, I):
    mse = K.mse_rows(c.out, c.target, B, I)
    return mse * c.loss_scale if c.loss_scale is not None else mse


def _loss_seed(c, g_mse):
    g = g_mse.float()
    return (g * c.loss_scale).contiguous() if c.loss_scale is not None else g.contiguous()


# ======================================================================================================
# GDMCF backbone
# ======================================================================================================
_GDMCF_PARAMS = ("emb_layer.weight", "emb_layer.bias", "in_layers.0.weight", "in_layers.0.bias", "in_layers2.0.weight",
                 "in_layers2.0.bias", "embedding_item.weight", "embedding_user.weight", "gcn_model.conv1.bias",
                 "gcn_model.conv1.lin.weight", "gcn_model.conv2.bias", "gcn_model.conv2.lin.weight", "sumW")


def _gdmcf_names(model):
    """Trained parameters of the configured model, in a fixed order: the base set + the encoder layers after the first."""
    have = dict(model.named_parameters())
    names = [n for n in _GDMCF_PARAMS if n in have]
    for branch in (0, 1):
        for name, _ in model._deep_layers(branch):
            names += [name + ".weight", name + ".bias"]
    return names


def _gdmcf_forward(model: DNNOneHotEmbeddingGCN, diff, x0, B, I, idx32, ts_disc, ts, inject) -> _Ctx:
    dev, d, T = x0.device, model.hidden, diff.steps
    inject = inject or {}
    p = model.drop.p if model.training else 0.0
    c = _Ctx()
    c.B, c.I, c.x0, c.ts, c.idx32 = B, I, x0, ts, idx32
    nt, gl = model.noise_type, model.gcn_layers
    # noising + dropout + operand cast, one pass each (gaussian_diffusion.py:849-870, DNN.py:1232-1233)
    _noised_input(diff, model, c, x0, B, I, ts, inject, p)
    c.A2 = torch.empty(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device=dev)
    kxu = inject.get("keep_xU")
    K.onehot_noise(x0, B, I, c.A2, ts=ts_disc, discrete=float(diff.discrete), dropout_p=p, u_keep=inject.get("u_keep"),
                   u_drop=kxu.float().contiguous() if kxu is not None else None, seed=diff.seed, offset=diff._offset(), epoch=diff._epoch)
    # user tower buffers: always with lo parts (nt_xent needs split precision)
    c.hc_f32 = torch.empty(B, 3 * d, dtype=torch.float32, device=dev)
    c.hc = Bf16Mat.empty(B, 3 * d, dev, True)
    bufs = dict(hc_f32=c.hc_f32, hc=c.hc, S=None)
    c.bufs = bufs  # deep encoders (dims with more than one entry) leave their layer activations in bufs["acts"]
    # encoder inputs: noise_type 1 feeds columns of the interleaved one-hot matrix to the continuous encoder (DNN.py:1236),
    # noise_type 2 feeds [x, x] to the one-hot encoder (:1246); they are also the wgrad operands of the backward pass
    c.enc1_in = model._x_branch_operand(c.A1, c.A2, B)
    model._encode_x(bufs, c.enc1_in, B, ts, 0, T)
    if nt == 2:
        xx = Bf16Mat.empty(B, 2 * I, dev, model._lo)
        for src, dst in ((c.A1.hi, xx.hi), (c.A1.lo, xx.lo)):
            if src is not None:
                dst[:, :I].copy_(src[:, :I])
                dst[:, I:2 * I].copy_(src[:, :I])
        c.enc2_in = xx
        model._encode_onehot_dense(bufs, xx.hi, B, ts, 0, T, to_S=False, lo=xx.lo)
    else:
        c.enc2_in = Bf16Mat(c.A2, None, B, 2 * I)
        model._encode_onehot_dense(bufs, c.A2, B, ts, 0, T, to_S=False)
    f32, hi, lo = model._seg(bufs, 2)
    K.gather_rows(model.embedding_user.weight.detach(), idx32, B, d, out_f32=f32, out_bf16=hi, out_bf16_lo=lo)
    c.closs_rows = torch.zeros(B, dtype=torch.float32, device=dev)
    c.S = None
    if nt == 0:  # noise_type != 0 multiplies the contrastive loss by zero (DNN.py:1258-1259)
        # nt_xent logits (DNN.py:488): S_raw = h h_U^T, split precision
        h_op = Bf16Mat(c.hc.hi[:, :d], c.hc.lo[:, :d], B, d)
        hu_op = Bf16Mat(c.hc.hi[:, d:2 * d], c.hc.lo[:, d:2 * d], B, d)
        c.S = torch.empty(B, K.round_up(B, 4), dtype=torch.float32, device=dev)
        _mm3(h_op, hu_op, B, B, d, out_f32=c.S)
        K.ntxent_rows(c.S, B, loss_rows=c.closs_rows)
    # GCN on user rows, mix, norms (DNN.py:1274-1288)
    g = model.gcn_model
    c.g2 = torch.empty(B, 3 * d, dtype=torch.float32, device=dev)
    c.hcp_f32 = torch.empty(B, 3 * d, dtype=torch.float32, device=dev)
    c.hcp = Bf16Mat.empty(B, 3 * d, dev, model._lo)
    c.inv_u = torch.empty(B, dtype=torch.float32, device=dev)
    hc_in = c.hc if model._lo else Bf16Mat(c.hc.hi, None, B, 3 * d)
    if gl == 0:    # no GCN (:1278): hc * sumW + hc * (1 - sumW)
        c.g2 = c.hc_f32
        K.mix_rownorm(c.hc_f32, B, 3 * d, g=c.hc_f32, sumw=model.sumW.detach(), out_f32=c.hcp_f32, out=c.hcp, inv_norm=c.inv_u)
    elif gl == 1:  # a single 3d -> 3d convolution, no activation (:1095-1096)
        wc1 = model._weight_operand("gcn1", g.conv1.lin.weight)
        _mm_auto(model, hc_in, wc1, B, 3 * d, 3 * d, bias=g.conv1.bias.detach(), out_f32=c.g2)
        K.mix_rownorm(c.hc_f32, B, 3 * d, g=c.g2, sumw=model.sumW.detach(), out_f32=c.hcp_f32, out=c.hcp, inv_norm=c.inv_u)
    if gl != 2:
        e_op, c.inv_i = model._item_operands()
        c.out = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
        _mm_auto(model, c.hcp, e_op, B, I, 3 * d, row_scale=c.inv_u, col_scale=c.inv_i, out_f32=c.out)
        c.mse = _row_mse(c, B, I)
        return c
    wc1 = model._weight_operand("gcn1", g.conv1.lin.weight)
    wc2 = model._weight_operand("gcn2", g.conv2.lin.weight)
    H = g.conv1.lin.weight.shape[0]
    c.g1_f32 = torch.empty(B, H, dtype=torch.float32, device=dev)
    if model._fused_tower():
        # one launch for both linears, the mix and the norms; the fp32 copies feed the backward pass
        K.user_tower(hc_in, c.hc_f32, wc1, g.conv1.bias.detach(), wc2, g.conv2.bias.detach(), model.sumW.detach(), B,
                     out=c.hcp, inv_u=c.inv_u, g1_f32=c.g1_f32, g2_f32=c.g2, hcp_f32=c.hcp_f32)
    else:
        c.g1 = Bf16Mat.empty(B, H, dev, model._lo)
        _mm_auto(model, hc_in, wc1, B, H, 3 * d, act=K.ACT_RELU, bias=g.conv1.bias.detach(), out_f32=c.g1_f32,
                 out_bf16=c.g1.hi, out_bf16_lo=c.g1.lo)
        _mm_auto(model, c.g1, wc2, B, 3 * d, H, bias=g.conv2.bias.detach(), out_f32=c.g2)
        K.mix_rownorm(c.hc_f32, B, 3 * d, g=c.g2, sumw=model.sumW.detach(), out_f32=c.hcp_f32, out=c.hcp, inv_norm=c.inv_u)
    # cosine scorer (DNN.py:1304-1327) and per-row MSE (gaussian_diffusion.py:902)
    e_op, c.inv_i = model._item_operands()
    c.out = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
    _mm_auto(model, c.hcp, e_op, B, I, 3 * d, row_scale=c.inv_u, col_scale=c.inv_i, out_f32=c.out)
    c.mse = _row_mse(c, B, I)
    return c


def _gdmcf_backward(model: DNNOneHotEmbeddingGCN, diff, c: _Ctx, g_mse: torch.Tensor, g_closs: torch.Tensor):
    grads: Dict[str, torch.Tensor] = {}
    for stage in _gdmcf_backward_stages(model, diff, c, g_mse, g_closs):
        grads.update(stage)
    return grads


def _gdmcf_backward_stages(model: DNNOneHotEmbeddingGCN, diff, c: _Ctx, g_mse: torch.Tensor, g_closs: torch.Tensor,
                           defer_item_norm: bool = False, sparse_user_grad: bool = False):
    """Generator over the backward pass: yields {parameter name: gradient} as soon as a group is final, so that a
    data-parallel caller can start its all-reduce while the rest of the backward runs. Stage 1: the item table
    (412 MB at the Yelp shape, needs only dL/d(out) and the user tower); stage 2: sumW, GCN linears, user table;
    stage 3: the first-layer weights."""
    B, I, d, dev, T = c.B, c.I, model.hidden, c.x0.device, diff.steps
    d3, e = 3 * d, model.time_emb_dim
    lo = model._lo
    P = dict(model.named_parameters())
    grads: Dict[str, torch.Tensor] = {}
    ones1 = _ones(model, 1, dev)
    Bp = K.round_up(B, 64)
    # ---- dL/d out, fused with operand production and the norm-term reductions
    Gs = Bf16Mat.empty(B, I, dev, lo, zero=False)  # loss_grad writes every [B, I] element; only the K tail needs zeros
    if Gs.ld > I:
        Gs.hi[:, I:].zero_()
        if Gs.lo is not None:
            Gs.lo[:, I:].zero_()
    # data-parallel engine, bf16 mode: the item table's gradient is exchanged as its factors (engine.StepEngine); Gs^T is
    # then produced straight into the engine's persistent send buffer and the local contraction is skipped
    fsend = getattr(model, "_item_factor_send", None) if (defer_item_norm and not lo) else None
    GsT = Bf16Mat(fsend[0], None, I, B) if fsend is not None else Bf16Mat.empty(I, B, dev, lo, zero=False)
    n_cb = (I + 31) // 32
    colsum = torch.empty(I, dtype=torch.float32, device=dev)
    rowpart = torch.empty(n_cb, B, dtype=torch.float32, device=dev)
    K.loss_grad(c.out, c.target, _loss_seed(c, g_mse), B, I, Gs, GT=GsT, row_scale=c.inv_u, col_scale=c.inv_i,
                with_out=True, colsum=colsum, rowpart=rowpart)
    # ---- d E = Gs^T hc' - E * ri^2 * c_i   (written straight into the parameter's gradient)
    hcpT = K.cast_bf16_transpose(c.hcp_f32, with_lo=lo)  # [3d, B]
    gE = _grad_buffer(P["embedding_item.weight"], model, "embedding_item.weight")
    coef_i = -(c.inv_i * c.inv_i) * colsum
    if fsend is not None:
        assert fsend[1].shape == hcpT.hi.shape
        fsend[1].copy_(hcpT.hi)
        model._item_grad_rowcoef = coef_i
        yield {"embedding_item.weight": None}
    elif defer_item_norm:
        # engine path: the row-wise norm term -E_i * ri^2 * c_i is applied by the optimizer pass, which reads E anyway
        # (FusedAdamW.update(row_coef=...)); the contraction then writes 412 MB instead of reading and writing it
        _mm_auto(model, GsT, hcpT, I, d3, B, out_f32=gE)
        model._item_grad_rowcoef = coef_i
    else:
        _mm_auto(model, GsT, hcpT, I, d3, B, out_f32=gE, row_t=_arange32(model, I, dev),
                 c1=_ones(model, I, dev), c2=coef_i, xt=P["embedding_item.weight"].detach())
        model._item_grad_rowcoef = None
    if fsend is None:
        yield {"embedding_item.weight": gE}
    # ---- cosine backward w.r.t. the user tower: d hc' = Gs E - hc' * ru^2 * r_b
    r_b = K.colsum_f32(rowpart, n_cb, B)
    eT = model._weight_operand("E", model.embedding_item.weight, transpose=True)  # [3d, I]
    d_hcp = torch.empty(B, d3, dtype=torch.float32, device=dev)
    coef_u = -(c.inv_u * c.inv_u) * r_b
    _mm_auto(model, Gs, eT, B, d3, I, out_f32=d_hcp, row_t=_arange32(model, B, dev),
             c1=_ones(model, B, dev), c2=coef_u, xt=c.hcp_f32)
    # ---- sumW mix backward
    d_hc = torch.empty(B, d3, dtype=torch.float32, device=dev)
    d_g2 = torch.empty(B, d3, dtype=torch.float32, device=dev)
    dw_rows = torch.empty(B, dtype=torch.float32, device=dev)
    K.mix_backward(d_hcp, c.hc_f32, c.g2, model.sumW.detach(), d_hc, d_g2, dw_rows, B, d3)
    grads["sumW"] = dw_rows.sum()
    # ---- GCN (user rows) backward
    g = model.gcn_model
    gl = model.gcn_layers
    d_hc_tot = torch.empty(B, d3, dtype=torch.float32, device=dev)
    if gl == 0:    # hc' = hc * sumW + hc * (1 - sumW): both mix branches lead straight back to hc
        K.ew_binary(K.EW_AXPBY, d_hc, d_g2, B, d3, alpha=1.0, beta=1.0, out_f32=d_hc_tot)
    elif gl == 1:  # g2 = hc W1^T + b1
        d_g2_op = K.cast_bf16(d_g2, with_lo=lo)
        d_g2T = K.cast_bf16_transpose(d_g2, with_lo=lo)      # [3d, B]
        hcT = K.cast_bf16_transpose(c.hc_f32, with_lo=lo)    # [3d, B]
        gW1 = torch.empty_like(P["gcn_model.conv1.lin.weight"])
        _mm_auto(model, d_g2T, hcT, d3, d3, B, out_f32=gW1)
        grads["gcn_model.conv1.lin.weight"] = gW1
        grads["gcn_model.conv1.bias"] = K.colsum_f32(d_g2, B, d3)
        wc1T = model._weight_operand("gcn1", g.conv1.lin.weight, transpose=True)  # [3d_in, 3d_out]
        _mm_auto(model, d_g2_op, wc1T, B, d3, d3, out_f32=d_hc_tot, c1=ones1, c2=ones1, xt=d_hc, t_const=0)
    if gl != 2:
        gU_rows = d_hc_tot[:, 2 * d:]
        model._user_grad_rows = (c.idx32, gU_rows)
        if not sparse_user_grad:
            gU = torch.zeros_like(P["embedding_user.weight"])
            K.scatter_rows_add(gU_rows, c.idx32, gU, B, d)
            grads["embedding_user.weight"] = gU
        yield grads
        yield _first_layer_grads(model, diff, c, d_hc_tot, g_closs, P)
        return
    d_g2_op = K.cast_bf16(d_g2, with_lo=lo)
    d_g2T = K.cast_bf16_transpose(d_g2, with_lo=lo)      # [3d, B]
    g1T = K.cast_bf16_transpose(c.g1_f32, with_lo=lo)    # [512, B]
    gW2 = torch.empty_like(P["gcn_model.conv2.lin.weight"])
    _mm_auto(model, d_g2T, g1T, d3, 512, B, out_f32=gW2)
    grads["gcn_model.conv2.lin.weight"] = gW2
    grads["gcn_model.conv2.bias"] = K.colsum_f32(d_g2, B, d3)
    wc2T = model._weight_operand("gcn2", g.conv2.lin.weight, transpose=True)  # [512, 3d]
    d_g1 = torch.empty(B, 512, dtype=torch.float32, device=dev)
    _mm_auto(model, d_g2_op, wc2T, B, 512, d3, out_f32=d_g1)
    d_pre1 = torch.empty(B, 512, dtype=torch.float32, device=dev)
    d_pre1_op = Bf16Mat.empty(B, 512, dev, lo, zero=True)
    K.ew_binary(K.EW_RELU_BWD, d_g1, c.g1_f32, B, 512, out_f32=d_pre1, out_bf16=d_pre1_op.hi, out_bf16_lo=d_pre1_op.lo)
    d_pre1T = K.cast_bf16_transpose(d_pre1, with_lo=lo)  # [512, B]
    hcT = K.cast_bf16_transpose(c.hc_f32, with_lo=lo)    # [3d, B]
    gW1 = torch.empty_like(P["gcn_model.conv1.lin.weight"])
    _mm_auto(model, d_pre1T, hcT, 512, d3, B, out_f32=gW1)
    grads["gcn_model.conv1.lin.weight"] = gW1
    grads["gcn_model.conv1.bias"] = K.colsum_f32(d_pre1, B, 512)
    wc1T = model._weight_operand("gcn1", g.conv1.lin.weight, transpose=True)  # [3d, 512]
    _mm_auto(model, d_pre1_op, wc1T, B, d3, 512, out_f32=d_hc_tot, c1=ones1, c2=ones1, xt=d_hc, t_const=0)
    # ---- user embedding rows
    # the rows of the user table that received a gradient (sparse exchange / row-sparse optimizer instead of a dense table)
    model._user_grad_rows = (c.idx32, d_hc_tot[:, 2 * d:])
    if not sparse_user_grad:
        gU = torch.zeros_like(P["embedding_user.weight"])
        K.scatter_rows_add(d_hc_tot[:, 2 * d:], c.idx32, gU, B, d)
        grads["embedding_user.weight"] = gU
    yield grads  # stage 2: sumW, the GCN linears, the user table — small messages, final before the two big wgrads
    yield _first_layer_grads(model, diff, c, d_hc_tot, g_closs, P)  # stage 3: the two first-layer weights (+ biases, emb_layer)


def _deep_backward(model, c: _Ctx, branch: int, dh_out, grads, B, dev):
    """Backward through the encoder layers after the first: dh_out = dL/d(last layer's tanh output) [B, hidden]. Adds the
    layers' weight / bias gradients to `grads`; returns (dL/d a1, a1) with a1 = the first layer's tanh output [B, d1]."""
    lo = model._lo
    acts = c.bufs["acts"][branch]            # acts[i] = (f32, operand) INPUT of deep layer i; acts[0] = first-layer output
    layers = model._deep_layers(branch)
    seg = c.hc_f32[:, branch * model.hidden:(branch + 1) * model.hidden]
    dh = dh_out
    for li in reversed(range(len(layers))):
        name, layer = layers[li]
        n_out, n_in = layer.weight.shape
        a_out = seg if li == len(layers) - 1 else acts[li + 1][0]
        a_in_f32 = acts[li][0]
        dpre = torch.empty(B, n_out, dtype=torch.float32, device=dev)
        dpre_op = Bf16Mat.empty(B, n_out, dev, lo, zero=True)
        K.ew_binary(K.EW_TANH_BWD, dh, a_out, B, n_out, out_f32=dpre, out_bf16=dpre_op.hi, out_bf16_lo=dpre_op.lo)
        dpreT = K.cast_bf16_transpose(dpre, with_lo=lo)        # [n_out, B]
        a_inT = K.cast_bf16_transpose(a_in_f32, with_lo=lo)    # [n_in, B]
        gW = torch.empty_like(layer.weight)
        _mm_auto(model, dpreT, a_inT, n_out, n_in, B, out_f32=gW)
        grads[name + ".weight"] = gW
        grads[name + ".bias"] = K.colsum_f32(dpre, B, n_out)
        wT = model._weight_operand(name, layer.weight, transpose=True)   # [n_in, n_out]
        dh = torch.empty(B, n_in, dtype=torch.float32, device=dev)
        _mm_auto(model, dpre_op, wT, B, n_in, n_out, out_f32=dh)
    return dh, acts[0][0]


def _first_layer_grads(model, diff, c: _Ctx, d_hc_tot, g_closs, P):
    """Stage 3 of the GDMCF backward: contrastive-loss gradient into h / h_U, tanh backward, the two first-layer weight
    gradients and the time-embedding layer."""
    B, I, d, dev, T = c.B, c.I, model.hidden, c.x0.device, diff.steps
    e = model.time_emb_dim
    lo = model._lo
    ones1 = _ones(model, 1, dev)
    grads: Dict[str, torch.Tensor] = {}
    if c.S is not None:
        # ---- nt_xent backward (DNN.py:479-508) into h and h_U
        dS = torch.empty(B, K.round_up(B, 4), dtype=torch.float32, device=dev)
        K.ntxent_rows(c.S, B, dscale=g_closs.float().reshape(1), dS=dS)
        dS_op = K.cast_bf16(dS[:, :B], with_lo=True)
        dST_op = K.cast_bf16_transpose(dS[:, :B], with_lo=True)
        hT = K.cast_bf16_transpose(c.hc_f32[:, :d], with_lo=True)          # [d, B]
        hUT = K.cast_bf16_transpose(c.hc_f32[:, d:2 * d], with_lo=True)    # [d, B]
        dh_tot = torch.empty(B, d, dtype=torch.float32, device=dev)
        dhU_tot = torch.empty(B, d, dtype=torch.float32, device=dev)
        _mm3(dS_op, hUT, B, d, B, out_f32=dh_tot, c1=ones1, c2=ones1, xt=d_hc_tot[:, :d], t_const=0)
        _mm3(dST_op, hT, B, d, B, out_f32=dhU_tot, c1=ones1, c2=ones1, xt=d_hc_tot[:, d:2 * d], t_const=0)
    else:  # noise_type != 0: the contrastive loss is multiplied by zero (DNN.py:1258-1259)
        dh_tot, dhU_tot = d_hc_tot[:, :d], d_hc_tot[:, d:2 * d]
    # ---- deep encoders: back through the tanh layers after the first (DNN.py:1240-1242, :1249-1251)
    h1, hU1, d1 = c.hc_f32[:, :d], c.hc_f32[:, d:2 * d], d
    if model.deep:
        dh_tot, h1 = _deep_backward(model, c, 0, dh_tot, grads, B, dev)
        dhU_tot, hU1 = _deep_backward(model, c, 1, dhU_tot, grads, B, dev)
        d = d1 = model.d1  # from here on d is the width of the first layers' outputs
    # ---- tanh backward, first-layer weight gradients
    dh_pre = torch.empty(B, d, dtype=torch.float32, device=dev)
    dhU_pre = torch.empty(B, d, dtype=torch.float32, device=dev)
    K.ew_binary(K.EW_TANH_BWD, dh_tot, h1, B, d, out_f32=dh_pre)
    K.ew_binary(K.EW_TANH_BWD, dhU_tot, hU1, B, d, out_f32=dhU_pre)
    dh_preT = K.cast_bf16_transpose(dh_pre, with_lo=lo)     # [d, B]
    dhU_preT = K.cast_bf16_transpose(dhU_pre, with_lo=lo)
    emb_table = model._emb_table(T)   # [T, e], produced with the forward's bias tables (same weights)
    emb_rows = torch.empty(B, e, dtype=torch.float32, device=dev)
    temb_rows = torch.empty(B, e, dtype=torch.float32, device=dev)
    K.gather_rows(emb_table, c.ts, B, e, out_f32=emb_rows)
    K.gather_rows(_temb_table(model, T, dev), c.ts, B, e, out_f32=temb_rows)
    A1T = _input_T_with_time_rows(model, c.enc1_in, emb_rows, B, I, e, dev)        # [I + e, B]
    A2T = _input_T_with_time_rows(model, c.enc2_in, emb_rows, B, 2 * I, e, dev)    # [2I + e, B]
    d_emb = torch.empty(B, e, dtype=torch.float32, device=dev)
    for name, dpre, dpreT, AT, n_in, first in (("in_layers.0", dh_pre, dh_preT, A1T, I, True),
                                              ("in_layers2.0", dhU_pre, dhU_preT, A2T, 2 * I, False)):
        W = P[name + ".weight"]
        gW = _grad_buffer(W, model, name + ".weight")
        if lo and AT.lo is None:
            # fp32 mode, one-hot operand (exact in bf16, no lo part): its time rows would be rounded to bf16, so the e
            # time-embedding columns keep their own fp32 product
            _mm_auto(model, dpreT, AT, d, n_in, B, out_f32=gW)
            K.sgemm_small(dpre, emb_rows, gW[:, n_in:], d, e, B, trans_a=True)
        else:
            _mm_auto(model, dpreT, AT, d, n_in + e, B, out_f32=gW)    # all columns: inputs [0, n_in) + time embedding
        grads[name + ".weight"] = gW
        grads[name + ".bias"] = K.colsum_f32(dpre, B, d)
        K.sgemm_small(dpre, _time_cols(model, name, W, n_in), d_emb, B, e, d, beta=0.0 if first else 1.0)
    gWe = torch.empty_like(P["emb_layer.weight"])
    K.sgemm_small(d_emb, temb_rows, gWe, e, e, B, trans_a=True)
    grads["emb_layer.weight"] = gWe
    grads["emb_layer.bias"] = K.colsum_f32(d_emb, B, e)
    return grads


class _GdmcfTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, diff, x0, B, I, idx32, ts_disc, ts, inject, *params):
        c = _gdmcf_forward(model, diff, x0, B, I, idx32, ts_disc, ts, inject)
        ctx.c, ctx.model, ctx.diff = c, model, diff
        ctx.names = _gdmcf_names(model)
        closs = c.closs_rows.mean()
        ctx.mark_non_differentiable(c.out)
        return c.mse, closs, c.out

    @staticmethod
    def backward(ctx, g_mse, g_closs, g_out):
        grads = _gdmcf_backward(ctx.model, ctx.diff, ctx.c, g_mse, g_closs)
        ctx.c = None
        return (None,) * 9 + _hand_over(ctx.model, grads, ctx.names)


# ======================================================================================================
# DNN backbone
# ======================================================================================================
_DNN_PARAMS = ("emb_layer.weight", "emb_layer.bias", "in_layers.0.weight", "in_layers.0.bias", "out_layers.0.weight",
               "out_layers.0.bias")


def _dnn_forward(model: DNN, diff, x0, B, I, ts, inject) -> _Ctx:
    dev, d, T = x0.device, model.hidden, diff.steps
    inject = inject or {}
    p = model.drop.p if model.training else 0.0
    c = _Ctx()
    c.B, c.I, c.x0, c.ts = B, I, x0, ts
    _noised_input(diff, model, c, x0, B, I, ts, inject, p)
    c.h_f32 = torch.empty(B, d, dtype=torch.float32, device=dev)
    c.h = Bf16Mat.empty(B, d, dev, model._lo)
    model._encode(c.A1, B, ts, 0, T, c.h, h_f32=c.h_f32)
    c.out = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
    c.acts = []  # (fp32, operand) outputs of the middle layers (dims with more than one entry)
    model._decode(model._mid_forward(c.h, B, acts=c.acts), B, c.out)
    c.mse = _row_mse(c, B, I)
    return c


def _dnn_names(model):
    """Trained parameters in a fixed order: emb_layer, first layer, the middle layers, the last out_layer."""
    names = ["emb_layer.weight", "emb_layer.bias", "in_layers.0.weight", "in_layers.0.bias"]
    for name, _ in model._middle():
        names += [name + ".weight", name + ".bias"]
    return names + [model._dec_name + ".weight", model._dec_name + ".bias"]


def _dnn_backward(model: DNN, diff, c: _Ctx, g_mse: torch.Tensor):
    B, I, d, dev, T, e = c.B, c.I, model.hidden, c.x0.device, diff.steps, model.time_emb_dim
    lo = model._lo
    P = dict(model.named_parameters())
    grads: Dict[str, torch.Tensor] = {}
    Bp = K.round_up(B, 64)
    G = Bf16Mat.empty(B, I, dev, lo, zero=True)
    GT = Bf16Mat.empty(I, B, dev, lo, zero=False)
    colsum = torch.empty(I, dtype=torch.float32, device=dev)
    K.loss_grad(c.out, c.target, _loss_seed(c, g_mse), B, I, G, GT=GT, with_out=False, colsum=colsum)
    dec, dd = model._dec_name, model.d_dec
    dec_in = c.acts[-1][0] if c.acts else c.h_f32   # the last out_layer's input [B, dd]
    grads[dec + ".bias"] = colsum
    # d W_out [I, dd] = G^T h
    hT = K.cast_bf16_transpose(dec_in, with_lo=lo)  # [dd, B]
    gWo = torch.empty_like(P[dec + ".weight"])
    _mm_auto(model, GT, hT, I, dd, B, out_f32=gWo)
    grads[dec + ".weight"] = gWo
    # d h = G W_out  (B operand = W_out^T [dd, I])
    woT = model._weight_operand("out0", model.out_layers[-1].weight, transpose=True)
    dh = torch.empty(B, dd, dtype=torch.float32, device=dev)
    _mm_auto(model, G, woT, B, dd, I, out_f32=dh)
    # middle layers (dims with more than one entry), last to first: tanh', weight / bias gradient, input gradient
    mid = model._middle()
    for li in reversed(range(len(mid))):
        name, layer = mid[li]
        n_out, n_in = layer.weight.shape
        a_in = c.acts[li - 1][0] if li > 0 else c.h_f32
        dpre = torch.empty(B, n_out, dtype=torch.float32, device=dev)
        dpre_op = Bf16Mat.empty(B, n_out, dev, lo, zero=True)
        K.ew_binary(K.EW_TANH_BWD, dh, c.acts[li][0], B, n_out, out_f32=dpre, out_bf16=dpre_op.hi, out_bf16_lo=dpre_op.lo)
        gWm = torch.empty_like(layer.weight)
        _mm_auto(model, K.cast_bf16_transpose(dpre, with_lo=lo), K.cast_bf16_transpose(a_in, with_lo=lo), n_out, n_in, B, out_f32=gWm)
        grads[name + ".weight"] = gWm
        grads[name + ".bias"] = K.colsum_f32(dpre, B, n_out)
        dh = torch.empty(B, n_in, dtype=torch.float32, device=dev)
        _mm_auto(model, dpre_op, model._weight_operand(name, layer.weight, transpose=True), B, n_in, n_out, out_f32=dh)
    dh_pre = torch.empty(B, d, dtype=torch.float32, device=dev)
    K.ew_binary(K.EW_TANH_BWD, dh, c.h_f32, B, d, out_f32=dh_pre)
    dh_preT = K.cast_bf16_transpose(dh_pre, with_lo=lo)
    W = P["in_layers.0.weight"]
    gW = _grad_buffer(W)
    emb_table = K.time_bias_table(model.emb_layer.weight.detach(), model.emb_layer.bias.detach(), W.detach(), I, None, T)[1]
    emb_rows = torch.empty(B, e, dtype=torch.float32, device=dev)
    temb_rows = torch.empty(B, e, dtype=torch.float32, device=dev)
    K.gather_rows(emb_table, c.ts, B, e, out_f32=emb_rows)
    K.gather_rows(_temb_table(model, T, dev), c.ts, B, e, out_f32=temb_rows)
    A1T = _input_T_with_time_rows(model, c.A1, emb_rows, B, I, e, dev)
    _mm_auto(model, dh_preT, A1T, d, I + e, B, out_f32=gW)
    grads["in_layers.0.weight"] = gW
    grads["in_layers.0.bias"] = K.colsum_f32(dh_pre, B, d)
    d_emb = torch.empty(B, e, dtype=torch.float32, device=dev)
    K.sgemm_small(dh_pre, _time_cols(model, "in_layers.0", W, I), d_emb, B, e, d)
    gWe = torch.empty_like(P["emb_layer.weight"])
    K.sgemm_small(d_emb, temb_rows, gWe, e, e, B, trans_a=True)
    grads["emb_layer.weight"] = gWe
    grads["emb_layer.bias"] = K.colsum_f32(d_emb, B, e)
    return grads


class _DnnTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, diff, x0, B, I, ts, inject, *params):
        c = _dnn_forward(model, diff, x0, B, I, ts, inject)
        ctx.c, ctx.model, ctx.diff = c, model, diff
        ctx.mark_non_differentiable(c.out)
        return c.mse, c.out

    @staticmethod
    def backward(ctx, g_mse, g_out):
        grads = _dnn_backward(ctx.model, ctx.diff, ctx.c, g_mse)
        ctx.c = None
        return (None,) * 7 + _hand_over(ctx.model, grads, _dnn_names(ctx.model))


# ======================================================================================================
# training_losses
# ======================================================================================================
def _prepare(diff, model, x_start, index, inject):
    """Shared front end of the autograd and the explicit training paths: dense start rows, timestep draws."""
    x0, _, _, users, B, I = diff._dense_start(x_start, want_op=False, lo=False)
    dev = x0.device
    diff._begin_step(dev)
    if index is None and users is not None:
        index = users
    if diff.noise_scale == 0.0:
        raise NotImplementedError("noise_scale == 0 (no diffusion) is outside the hot path")
    gdmcf = isinstance(model, DNNOneHotEmbeddingGCN)
    if not gdmcf and not isinstance(model, DNN):
        raise TypeError("training_losses needs a gdmcf_b200 denoiser (DNN / DNNOneHotEmbeddingGCN)")
    if gdmcf != bool(diff.CatOneHot and diff.indexIn):
        raise ValueError("diffusion flags (CatOneHot/indexIn) do not match the backbone")
    # two timestep draws like the reference (:845 shapes the discrete noise, :865 feeds the model and the weights)
    ts_disc = None
    if gdmcf:
        ts_disc = inject["ts_discrete"] if "ts_discrete" in inject else diff.sample_timesteps(B, dev, "importance")[0]
        ts_disc = ts_disc.to(dev).to(torch.int32)
    if "ts" in inject:
        ts = inject["ts"].to(dev).long()
        pt = diff._pt_for(ts)
    else:
        ts, pt = diff.sample_timesteps(B, dev, "importance")
    idx32 = None
    if gdmcf:
        assert index is not None, "DNNOneHotEmbeddingGCN needs the user index of every row"
        idx32 = index.to(dev).to(torch.int32)
    return x0, B, I, dev, gdmcf, idx32, ts_disc, ts, ts.to(torch.int32), pt


def _loss_weight(diff, ts, B, dev, reweight):
    if not reweight:
        return torch.ones(B, device=dev)
    if _is_eps(diff):  # gaussian_diffusion.py:924-926 (float64 schedule tensors; t == 0 divides by zero, then overwritten)
        ac, acp, betas = (getattr(diff, n).to(dev) for n in ("alphas_cumprod", "alphas_cumprod_prev", "betas"))
        weight = (1 - ac[ts]) / ((1 - acp[ts]) ** 2 * (1 - betas[ts]))
        return torch.where((ts == 0), 1.0, weight)
    weight = diff.SNR(ts - 1) - diff.SNR(ts)
    return torch.where((ts == 0), 1.0, weight)


def training_losses(diff, model, x_start, reweight=False, index=None, inject: Optional[dict] = None):
    """GaussianDiffusionDiscrete.training_losses (models/gaussian_diffusion.py:834-957). Returns {"loss": [B] f64}
    (+ "model_output", "mse", "closs" for tests). `inject` may carry ts_discrete / ts / noise / u_keep / keep_x /
    keep_xU to replace the in-kernel Philox draws."""
    inject = inject or {}
    x0, B, I, dev, gdmcf, idx32, ts_disc, ts, ts32, pt = _prepare(diff, model, x_start, index, inject)
    params = dict(model.named_parameters())
    if gdmcf:
        mse, closs, out = _GdmcfTrainFn.apply(model, diff, x0, B, I, idx32, ts_disc, ts32, inject,
                                              *[params[n] for n in _gdmcf_names(model)])
    else:
        mse, out = _DnnTrainFn.apply(model, diff, x0, B, I, ts32, inject, *[params[n] for n in _dnn_names(model)])
        closs = None
    terms = {}
    terms["loss"] = _loss_weight(diff, ts, B, dev, reweight) * mse
    diff._update_history(ts, terms["loss"])
    terms["loss"] = terms["loss"] / pt
    if closs is not None:
        terms["loss"] = terms["loss"] + closs * 0.1
    terms["model_output"], terms["mse"], terms["closs"] = out.detach()[:, :I], mse.detach(), closs.detach() if closs is not None else None
    return terms


@torch.no_grad()
def fused_train_stages(diff, model, x_start, reweight=False, index=None, inject: Optional[dict] = None,
                       defer_item_norm: bool = False, sparse_user_grad: bool = False):
    """`training_losses(...)["loss"].mean().backward()` without autograd, as a generator (used by engine.StepEngine):
    yields ("loss", mean loss [] f64) after the forward and loss bookkeeping, then ("grads", {parameter name: gradient})
    once per backward stage, in the order the gradients become final — a data-parallel caller starts the all-reduce of a
    stage while the generator computes the next one. Same kernels and arithmetic as the autograd path; the gradient
    seeds are d mean(loss) / d mse_b = weight_b / (pt_b * B) and d mean(loss) / d closs = 0.1.
    defer_item_norm: the item table's gradient is yielded WITHOUT its row-wise norm term; the caller must pass
    model._item_grad_rowcoef to FusedAdamW.update(row_coef=...) (and all-reduce it with the gradient)."""
    inject = inject or {}
    x0, B, I, dev, gdmcf, idx32, ts_disc, ts, ts32, pt = _prepare(diff, model, x_start, index, inject)
    if gdmcf:
        c = _gdmcf_forward(model, diff, x0, B, I, idx32, ts_disc, ts32, inject)
        closs = c.closs_rows.mean()
    else:
        c = _dnn_forward(model, diff, x0, B, I, ts32, inject)
        closs = None
    if (not _is_eps(diff) and c.mse.is_cuda and c.mse.dtype == torch.float32 and pt.dtype == torch.float64
            and ts.dtype == torch.int64):
        # the per-row loss terms (weight, history entry, loss, backward seed) in one launch instead of ~20 tensor ops
        diff.alphas_cumprod = diff.alphas_cumprod.to(dev)
        hist, loss, g_mse = K.loss_terms(ts.contiguous(), pt.contiguous(), c.mse.contiguous(), diff.alphas_cumprod.contiguous(),
                                         closs.reshape(1) if closs is not None else None, reweight)
        diff._update_history(ts, hist)
    else:
        weight = _loss_weight(diff, ts, B, dev, reweight)
        loss = weight * c.mse
        diff._update_history(ts, loss)
        loss = loss / pt
        if closs is not None:
            loss = loss + closs * 0.1
        g_mse = (weight / pt / B).float()
    yield "loss", loss.mean()
    if gdmcf:
        g_closs = _vec_cache(model, ("g_closs", str(dev)), lambda: torch.full((), 0.1, dtype=torch.float32, device=dev))
        # sparse_user_grad: the user table's gradient is NOT materialised as a dense [n_user, d] tensor; the caller
        # consumes model._user_grad_rows = (user ids [B], gradient rows [B, d]) after the second stage
        for stage in _gdmcf_backward_stages(model, diff, c, g_mse, g_closs, defer_item_norm, sparse_user_grad):
            yield "grads", stage
    else:
        yield "grads", _dnn_backward(model, diff, c, g_mse)
