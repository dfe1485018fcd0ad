"""GaussianDiffusionDiscrete — host-side mirror of the reference's models/gaussian_diffusion.py:552-1135.

Same constructor, attributes (`indexIn`, `gcn`, `Lt_history`, `Lt_count`, schedule tensors) and methods
(`training_losses`, `p_sample`, `sample_timesteps`, `q_sample`, `apply_noise`, `p_mean_variance`,
`q_posterior_mean_variance`, `SNR`, `_extract_into_tensor`) as the reference; the arithmetic on [B, n_item]
tensors runs in libgdmcf_sm100.so. Extensions that the reference lacks (all optional, defaults keep the
reference behaviour): CSR batches instead of dense rows (`CsrBatch`), fused mask+top-K ranking (`rank`),
`precision`, injected random draws (`inject=`) for parity tests.

Both mean types (START_X / EPSILON) and sampling_noise are covered. The per-step random graph bookkeeping of p_sample
(gaussian_diffusion.py:710-729) cannot change any result (it only feeds GCN item rows the model discards)
and is only executed in the faithful-graph mode (`faithful_graph`, see sample_graph_step).
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from ..kernels import Bf16Mat


class ModelMeanType(enum.Enum):
    START_X = enum.auto()  # the model predicts x_0
    EPSILON = enum.auto()  # the model predicts epsilon


@dataclass
class CsrBatch:
    """A batch of users given as rows of a device CSR interaction matrix (replaces dense `batch` rows)."""
    rowptr: torch.Tensor  # int32 [n_user + 1]
    col: torch.Tensor     # int32 [nnz]
    users: torch.Tensor   # int32 [B]
    n_item: int

    @property
    def shape(self):
        return (self.users.numel(), self.n_item)


def betas_from_linear_variance(steps, variance, max_beta=0.999):
    """models/gaussian_diffusion.py:1138-1144."""
    alpha_bar = 1 - variance
    betas = [1 - alpha_bar[0]]
    for i in range(1, steps):
        betas.append(min(1 - alpha_bar[i] / alpha_bar[i - 1], max_beta))
    return np.array(betas)


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """models/gaussian_diffusion.py:1146-1162."""
    betas = []
    for i in range(num_diffusion_timesteps):
        t1, t2 = i / num_diffusion_timesteps, (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


def mean_flat(tensor):
    """models/gaussian_diffusion.py:1194-1198."""
    return tensor.mean(dim=list(range(1, len(tensor.shape))))


class GaussianDiffusionDiscrete(nn.Module):
    def __init__(self, mean_type, noise_schedule, noise_scale, noise_min, noise_max, steps, device,
                 history_num_per_term=10, beta_fixed=True, discrete=0.99, CatOneHot=False, epps=0.9995, args=None):
        self.args = args
        self.mean_type = mean_type
        self.noise_schedule = noise_schedule
        self.noise_scale = noise_scale
        self.noise_min = noise_min
        self.noise_max = noise_max
        self.steps = steps
        self.device = device
        self.discrete = discrete
        self.discrete_noise = True
        self.CatOneHot = CatOneHot
        self.history_num_per_term = history_num_per_term
        self.Lt_history = torch.zeros(steps, history_num_per_term, dtype=torch.float64).to(device)
        self.Lt_count = torch.zeros(steps, dtype=int).to(device)
        self.indexIn = False
        self.seed = 0
        self._calls = 0
        self._epoch = None
        self._importance_ready = False
        if mean_type not in (ModelMeanType.START_X, ModelMeanType.EPSILON):
            raise NotImplementedError(mean_type)
        if noise_scale != 0.0:
            self.betas = torch.tensor(self.get_betas(), dtype=torch.float64).to(self.device)
            if beta_fixed:
                self.betas[0] = 0.00001
            assert len(self.betas.shape) == 1, "betas must be 1-D"
            assert len(self.betas) == self.steps, "num of betas must equal to diffusion steps"
            assert (self.betas > 0).all() and (self.betas <= 1).all(), "betas out of range"
            self.calculate_for_diffusion()
        super(GaussianDiffusionDiscrete, self).__init__()
        epps = self.discrete  # the ctor argument `epps` is overwritten (gaussian_diffusion.py:589)
        self.u_x = torch.tensor([[epps, 1 - epps], [epps, 1 - epps]]).unsqueeze(0).to(device)
        self.u_x_eye = torch.eye(2).unsqueeze(0).to(device)

    # -- schedules (float64, as the reference) ---------------------------------------------------
    def get_betas(self):
        if self.noise_schedule in ("linear", "linear-var"):
            start = self.noise_scale * self.noise_min
            end = self.noise_scale * self.noise_max
            if self.noise_schedule == "linear":
                return np.linspace(start, end, self.steps, dtype=np.float64)
            return betas_from_linear_variance(self.steps, np.linspace(start, end, self.steps, dtype=np.float64))
        elif self.noise_schedule == "cosine":
            return betas_for_alpha_bar(self.steps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
        elif self.noise_schedule == "binomial":
            ts = np.arange(self.steps)
            return [1 / (self.steps - t + 1) for t in ts]
        raise NotImplementedError(f"unknown beta schedule: {self.noise_schedule}!")

    def calculate_for_diffusion(self):
        alphas = 1.0 - self.betas
        dev = self.device
        self.alphas_cumprod = torch.cumprod(alphas, axis=0).to(dev)
        self.alphas_cumprod_prev = torch.cat([torch.tensor([1.0]).to(dev), self.alphas_cumprod[:-1]]).to(dev)
        self.alphas_cumprod_next = torch.cat([self.alphas_cumprod[1:], torch.tensor([0.0]).to(dev)]).to(dev)
        assert self.alphas_cumprod_prev.shape == (self.steps,)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = torch.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(
            torch.cat([self.posterior_variance[1].unsqueeze(0), self.posterior_variance[1:]]))
        self.posterior_mean_coef1 = self.betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * torch.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        # fp32 device tables consumed by the kernels (the reference gathers the f64 value and casts, :1131)
        self._f32 = {k: getattr(self, k).float().contiguous() for k in (
            "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2")}
        # reverse-step coefficients x_{t-1} = ra[t] * model_output + rb[t] * x_t:
        #   START_X: pred_xstart = output                               -> ra = coef1, rb = coef2          (:1085-1086)
        #   EPSILON: pred_xstart = sr[t] * x_t - srm1[t] * output       -> ra = -coef1 * srm1, rb = coef1 * sr + coef2
        #            (_predict_xstart_from_eps :1106-1111 folded into q_posterior_mean_variance :1041-1050, in float64)
        if self.mean_type == ModelMeanType.EPSILON:
            ra = -self.posterior_mean_coef1 * self.sqrt_recipm1_alphas_cumprod
            rb = self.posterior_mean_coef1 * self.sqrt_recip_alphas_cumprod + self.posterior_mean_coef2
        else:
            ra, rb = self.posterior_mean_coef1, self.posterior_mean_coef2
        self._f32["reverse_a"], self._f32["reverse_b"] = ra.float().contiguous(), rb.float().contiguous()
        # the last reverse step returns the model output itself when ra[0] = 1, rb[0] = 0 (START_X: alphas_cumprod_prev[0] = 1)
        self._last_step_is_output = bool(float(ra[0].float()) == 1.0 and float(rb[0].float()) == 0.0)
        # sampling_noise (:745-750): x_{t-1} = mean + [t != 0] * exp(0.5 * log_variance[t]) * N(0, 1)
        sigma = torch.exp(0.5 * self.posterior_log_variance_clipped)
        sigma[0] = 0.0
        self._f32["noise_sigma"] = sigma.float().contiguous()
        self._f32["ones"] = torch.ones_like(self._f32["noise_sigma"])

    def get_Qt_bar(self, alpha_bar_t):
        alpha_bar_t = alpha_bar_t.unsqueeze(1).unsqueeze(1)
        return alpha_bar_t * self.u_x_eye + (1 - alpha_bar_t) * self.u_x

    def SNR(self, t):
        self.alphas_cumprod = self.alphas_cumprod.to(t.device)
        return self.alphas_cumprod[t] / (1 - self.alphas_cumprod[t])

    def _extract_into_tensor(self, arr, timesteps, broadcast_shape):
        arr = arr.to(timesteps.device)
        res = arr[timesteps].float()
        while len(res.shape) < len(broadcast_shape):
            res = res[..., None]
        return res.expand(broadcast_shape)

    def _begin_step(self, dev) -> None:
        """Advance the device-resident RNG epoch (one per training_losses / p_sample call). The 128-bit Philox counter of
        every in-kernel draw is low word = (call site << 40) + element, high word = (epoch << 8) | sub-stream, so a captured
        CUDA graph draws fresh numbers on each replay without any host-side state and the epoch (56 bits) can never run
        into the call-site field."""
        if self._epoch is None or self._epoch.device != torch.device(dev):
            self._epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        K.counter_add(self._epoch, 1)
        self._calls = 0

    def _offset(self) -> int:
        self._calls += 1
        assert self._calls < (1 << 20), "more than 2^20 RNG call sites inside one step"
        return (self._calls << 40) | (1 << 62)

    # -- inputs ----------------------------------------------------------------------------------
    def _dense_start(self, x_start, want_op: bool, lo: bool):
        """Returns (x0_f32 [B, ld4], x0_op or None, csr, users, B, I) for dense tensors or CsrBatch inputs."""
        if isinstance(x_start, CsrBatch):
            B, I = x_start.shape
            dev = x_start.col.device
            x0 = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
            op = Bf16Mat.empty(B, I, dev, lo, zero=False) if want_op else None
            K.densify_rows(x_start.rowptr, x_start.col, x_start.users, B, I, out_f32=x0, out_bf16=op.hi if op else None)
            if op is not None and op.lo is not None:
                op.lo.zero_()  # {0,1} is exact in bf16
            return x0, op, (x_start.rowptr, x_start.col), x_start.users, B, I
        K.require_cuda(x_start)
        B, I = x_start.shape
        x0 = x_start if (x_start.dtype == torch.float32 and x_start.stride(1) == 1) else x_start.float().contiguous()
        return x0, None, None, None, B, I

    # -- forward process -------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        """gaussian_diffusion.py:988-996 on the device (Philox noise when `noise` is None)."""
        K.require_cuda(x_start)
        B, I = x_start.shape
        if noise is not None:
            assert noise.shape == x_start.shape
            noise = noise.float().contiguous()
        out = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=x_start.device)
        scratch = Bf16Mat.empty(B, I, x_start.device, zero=False)
        self._begin_step(x_start.device)
        K.qsample_dropout(x_start.float(), B, I, scratch, row_t=t.to(torch.int32), sqrt_ab=self._f32["sqrt_alphas_cumprod"],
                          sqrt_1mab=self._f32["sqrt_one_minus_alphas_cumprod"], noise=noise, seed=self.seed,
                          offset=self._offset(), epoch=self._epoch, xt_out=out)
        return out[:, :I]

    def apply_noise(self, ts, x_start, x_base=None, u_keep=None):
        """gaussian_diffusion.py:770-831: x_start is the one-hot [B, I, 2] tensor; returns the sampled one-hot
        as int64 [B, I, 2]. (The fused training path never materialises this tensor; kept for API parity.)"""
        B, I, _ = x_start.shape
        cls = x_start[..., 1].float().contiguous()  # class index as {0,1}
        out = torch.zeros(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device=x_start.device)
        self._begin_step(x_start.device)
        K.onehot_noise(cls, B, I, out, ts=ts.to(torch.int32), discrete=float(self.discrete), u_keep=u_keep, seed=self.seed,
                       offset=self._offset(), epoch=self._epoch)
        kept = out[:, : 2 * I].reshape(B, I, 2) != 0
        # apply_noise itself returns one_hot(sample): the class flips when the true class was not kept
        onehot = x_start.bool()
        flipped = torch.stack([onehot[..., 1], onehot[..., 0]], dim=-1)
        any_kept = kept.any(dim=-1, keepdim=True)
        return torch.where(any_kept, onehot, flipped).long()

    # -- timestep sampling (gaussian_diffusion.py:959-986) ----------------------------------------
    def sample_timesteps(self, batch_size, device, method="uniform", uniform_prob=0.001):
        if method == "importance" and self.Lt_history.is_cuda:
            # one kernel, no host sync (the reference's `.all()` / multinomial checks synchronise every call):
            # uniform with pt = 1 until every Lt_count reaches history_num_per_term, importance sampling afterwards
            if self._epoch is None:
                self._begin_step(self.Lt_history.device)
            return K.sample_timesteps(self.Lt_history, self.Lt_count, batch_size, uniform_prob=uniform_prob, seed=self.seed,
                                      offset=self._offset(), epoch=self._epoch)
        if method == "importance":
            if not self._importance_ready:
                self._importance_ready = bool((self.Lt_count == self.history_num_per_term).all())
            if not self._importance_ready:
                return self.sample_timesteps(batch_size, device, method="uniform")
            Lt_sqrt = torch.sqrt(torch.mean(self.Lt_history ** 2, axis=-1))
            pt_all = Lt_sqrt / torch.sum(Lt_sqrt)
            pt_all *= 1 - uniform_prob
            pt_all += uniform_prob / len(pt_all)
            t = torch.multinomial(pt_all, num_samples=batch_size, replacement=True)
            pt = pt_all.gather(dim=0, index=t) * len(pt_all)
            return t, pt
        elif method == "uniform":
            t = torch.randint(0, self.steps, (batch_size,), device=device).long()
            pt = torch.ones_like(t).float()
            return t, pt
        raise ValueError

    def _pt_for(self, ts):
        if self.Lt_history.is_cuda:
            return K.sample_timesteps(self.Lt_history, self.Lt_count, ts.numel(), ts_in=ts.contiguous())[1]
        if not self._importance_ready:
            self._importance_ready = bool((self.Lt_count == self.history_num_per_term).all())
        if not self._importance_ready:
            return torch.ones_like(ts).float()
        Lt_sqrt = torch.sqrt(torch.mean(self.Lt_history ** 2, axis=-1))
        pt_all = Lt_sqrt / torch.sum(Lt_sqrt)
        pt_all = pt_all * (1 - 0.001) + 0.001 / len(pt_all)
        return pt_all.gather(0, ts) * len(pt_all)

    def _update_history(self, ts, loss):
        """gaussian_diffusion.py:935-949 without the per-sample Python loop / host syncs: for every t the new
        history is the last `history_num_per_term` entries of (old entries ++ this batch's losses for t, in batch
        order) — exactly what the sequential shift-and-append loop leaves behind."""
        H = self.history_num_per_term
        T = self.steps
        loss = loss.detach().to(torch.float64)
        if loss.is_cuda:  # one tiny kernel: thread t replays the batch in order (no host sync)
            K.lt_history_update(ts.contiguous(), loss.contiguous(), self.Lt_history, self.Lt_count)
            return
        # host (CPU tensors, used by the gloo tests): same result, vectorised
        onehot = torch.nn.functional.one_hot(ts, T)                      # [B, T]
        rank = torch.cumsum(onehot, 0) - onehot                          # occurrences of t before row b
        n_t = onehot.sum(0)                                              # [T]
        pos = self.Lt_count[ts] + (rank * onehot).sum(1)                 # slot in (old ++ new)
        total = self.Lt_count + n_t
        shift = (total - H).clamp(min=0)                                 # entries that fall off the front
        old_idx = torch.arange(H, device=ts.device)[None, :].expand(T, H)
        old_keep = (old_idx < self.Lt_count[:, None]) & (old_idx >= shift[:, None])
        new_hist = torch.zeros_like(self.Lt_history)
        dst_old = (old_idx - shift[:, None]).clamp(min=0)
        new_hist.scatter_add_(1, dst_old, torch.where(old_keep, self.Lt_history, torch.zeros_like(self.Lt_history)))
        dst_new = pos - shift[ts]
        ok = (dst_new >= 0).to(loss.dtype)  # entries pushed out by later ones of the same batch contribute 0
        new_hist.index_put_((ts, dst_new.clamp(min=0)), loss * ok, accumulate=True)
        self.Lt_history = new_hist
        self.Lt_count = torch.minimum(total, torch.full_like(total, H))

    # -- training (gaussian_diffusion.py:834-957) -------------------------------------------------
    def training_losses(self, model, x_start, reweight=False, index=None, inject=None):
        from ..train_step import training_losses as _tl
        return _tl(self, model, x_start, reweight, index, inject)

    # -- sampling (gaussian_diffusion.py:668-768) --------------------------------------------------
    @torch.no_grad()
    def p_sample(self, model, x_start, steps, sampling_noise=False, index=None, inject=None, _raw=False):
        """Reverse process over all `self.steps` timesteps starting from x_start (steps == 0) or its
        q_sample at t = steps-1. Returns fp32 [B, n_item]. x_start: dense fp32 CUDA tensor or CsrBatch."""
        assert steps <= self.steps, "Too much steps in inference."
        if not hasattr(model, "reverse_loop"):
            raise TypeError("p_sample needs a gdmcf_b200 denoiser (DNN / DNNOneHotEmbeddingGCN); there is no generic torch path")
        lo = getattr(model, "_lo", False)
        x0, x0_op, csr, users, B, I = self._dense_start(x_start, want_op=(steps == 0), lo=lo)
        dev = x0.device
        if steps != 0 or sampling_noise:
            self._begin_step(dev)
        gdmcf = self.CatOneHot and self.indexIn
        if self.CatOneHot and not self.indexIn:
            raise NotImplementedError("CatOneHot without indexIn selects backbones outside the hot path")
        if index is None and users is not None:
            index = users
        idx32 = None
        if gdmcf:
            assert index is not None, "DNNOneHotEmbeddingGCN needs the user index of every row"
            idx32 = index.to(dev).to(torch.int32)
        xu_op = None
        x_t, x_op = x0, x0_op
        if steps != 0:
            # x_t = q_sample(x_start, steps-1); x_tU = apply_noise(...) & one_hot(x_start)  (:671-692)
            noise = inject.get("noise") if inject else None
            x_t = torch.empty(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
            x_op = Bf16Mat.empty(B, I, dev, lo, zero=False)
            K.qsample_dropout(x0, B, I, x_op, t_const=steps - 1, sqrt_ab=self._f32["sqrt_alphas_cumprod"],
                              sqrt_1mab=self._f32["sqrt_one_minus_alphas_cumprod"], noise=noise, seed=self.seed,
                              offset=self._offset(), epoch=self._epoch, xt_out=x_t)
            if gdmcf:
                xu_op = torch.zeros(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device=dev)
                ts = torch.full((B,), steps - 1, dtype=torch.int32, device=dev)
                K.onehot_noise(x0, B, I, xu_op, ts=ts, discrete=float(self.discrete),
                               u_keep=inject.get("u_keep") if inject else None, seed=self.seed, offset=self._offset(),
                               epoch=self._epoch)
        elif gdmcf and csr is None:
            # dense input, x_tU = one_hot(x0): build the interleaved one-hot operand once
            xu_op = torch.zeros(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device=dev)
            K.onehot_noise(x0, B, I, xu_op)
        if x_t.shape[1] != K.round_up(I, 4) or x_t.stride(0) != K.round_up(I, 4):
            xp = torch.zeros(B, K.round_up(I, 4), dtype=torch.float32, device=dev)
            xp[:, :I] = x_t[:, :I]
            x_t = xp
        if self.noise_scale == 0.0:
            raise NotImplementedError("noise_scale == 0 (no diffusion) is outside the hot path")
        c1, c2 = self._f32["reverse_a"], self._f32["reverse_b"]
        noise_hook = None
        if sampling_noise:
            draws = inject.get("sampling_noise") if inject else None  # optional: one [B, I] tensor per step, t = T-1 .. 0

            def noise_hook(x_f32, x_op_, t):  # x_{t-1} = mean + sigma[t] * z, t != 0; refreshes the bf16 operand too
                if t == 0:
                    return
                z = draws[self.steps - 1 - t].float().contiguous() if draws is not None else None
                K.qsample_dropout(x_f32, B, I, x_op_, t_const=t, sqrt_ab=self._f32["ones"], sqrt_1mab=self._f32["noise_sigma"],
                                  noise=z, seed=self.seed, offset=self._offset(), epoch=self._epoch, xt_out=x_f32)
        graph_hook = None
        if gdmcf and getattr(model, "faithful_graph", False):
            # faithful graph mode: the per-step random edge set of the reference (:710-729) — apply_noise on the accumulated
            # graph, degree-guided user draw (x_degree / x_degree.max() is a batch-global reduction, :711-712), OR-accumulate
            self._graph_state = torch.zeros(B, I, dtype=torch.uint8, device=dev)
            deg = x0[:, :I].sum(dim=1)
            deg_frac = (deg / deg.max()).float().contiguous()
            guided = bool(getattr(self.args, "user_guided", 1)) if self.args is not None else True
            g_inj = inject.get("graph_draws") if inject else None  # optional [(u_entry [B, I], u_user [B])] per step
            if self._epoch is None:
                self._begin_step(dev)

            def graph_hook(t):
                u_e, u_u = g_inj[self.steps - 1 - t] if g_inj is not None else (None, None)
                K.graph_noise_step(self._graph_state, t, B, deg_frac=deg_frac, discrete=float(self.discrete), user_guided=guided,
                                   seed=self.seed, offset=self._offset(), epoch=self._epoch, u_entry=u_e, u_user=u_u)
                return self._graph_state
        if gdmcf:
            if getattr(model, "needs_dense_onehot", False) and xu_op is None:
                xu_op = torch.zeros(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device=dev)
                K.onehot_noise(x0, B, I, xu_op)
            if (noise_hook is None and graph_hook is None and self._last_step_is_output and hasattr(model, "can_project")
                    and model.can_project()):
                # the recurrence carried in the encoder's pre-activation space: one catalogue-wide scorer instead of T
                out = model.reverse_loop_projected(x_t, B, idx32, self.steps, c1, c2, x0_op=x_op, csr=csr, users=users, xu_op=xu_op)
            else:
                out = model.reverse_loop(x_t, B, idx32, self.steps, c1, c2, x0_op=x_op, csr=csr, users=users, xu_op=xu_op,
                                         noise_hook=noise_hook, graph_hook=graph_hook)
        else:
            out = model.reverse_loop(x_t, B, None, self.steps, c1, c2, x0_op=x_op, noise_hook=noise_hook)
        # the loop's result lives in a cached ping-pong buffer that the next call overwrites: the public API hands back a
        # fresh tensor like the reference; rank() consumes the buffer in place (_raw)
        return out if _raw else out[:, :I].clone()

    @torch.no_grad()
    def rank(self, model, x_start, k, hist=None, hist2=None, steps=0, index=None, with_values=False, sampling_noise=False):
        """Fused evaluate step (main.py:288-304): p_sample -> history mask -> top-k, all on the device.
        hist / hist2: (rowptr, col) device CSR of the items to mask, indexed by the batch's user ids."""
        out = self.p_sample(model, x_start, steps, sampling_noise, index=index, _raw=True)
        B, I = x_start.shape
        users = x_start.users if isinstance(x_start, CsrBatch) else (index.to(out.device).to(torch.int32) if index is not None else None)
        return K.mask_topk(out, B, I, k, users=users, hist=hist, hist2=hist2, with_values=with_values)

    # -- API parity helpers ------------------------------------------------------------------------
    def q_posterior_mean_variance(self, x_start, x_t, t):
        assert x_start.shape == x_t.shape
        posterior_mean = (self._extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start
                          + self._extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        posterior_variance = self._extract_into_tensor(self.posterior_variance, t, x_t.shape)
        posterior_log_variance_clipped = self._extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    def _predict_xstart_from_eps(self, x_t, t, eps):
        """gaussian_diffusion.py:1106-1111."""
        assert x_t.shape == eps.shape
        return (self._extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - self._extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def p_mean_variance(self, model, x, t, x_tU=None, index=None, graph=None):
        """gaussian_diffusion.py:1063-1103 (START_X): one denoiser call + posterior mean (API parity; p_sample
        uses the fused loop instead)."""
        B, C = x.shape[:2]
        assert t.shape == (B,)
        if self.CatOneHot:
            model_output = model(x, t, x_tU, index=index, graph=graph)
        else:
            model_output = model(x, t)
        model_variance = self._extract_into_tensor(self.posterior_variance, t, x.shape)
        model_log_variance = self._extract_into_tensor(self.posterior_log_variance_clipped, t, x.shape)
        if self.mean_type == ModelMeanType.START_X:
            pred_xstart = model_output
        else:
            pred_xstart = self._predict_xstart_from_eps(x, t, eps=model_output)
        model_mean, _, _ = self.q_posterior_mean_variance(x_start=pred_xstart, x_t=x, t=t)
        return {"mean": model_mean, "variance": model_variance, "log_variance": model_log_variance,
                "pred_xstart": pred_xstart}
