"""Fused AdamW (torch.optim.AdamW semantics, main.py:258,351) on the gdmcf_adamw_fused kernel: one launch per
parameter, p/m/v/grad streamed once (28 B per element). Parameters without a gradient are skipped exactly like
torch (DNNOneHotEmbeddingGCN.out_layers never receives one). `grad_scale` folds the 1/world_size of a
data-parallel gradient sum into the update."""
from __future__ import annotations

import torch

from . import kernels as K


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, modules=(), capturable=False,
                 fuse_refresh=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._modules = list(modules)  # engine modules whose bf16 operand caches must be invalidated
        # capturable: the step count lives on the device (one counter shared by all parameters, advanced by a kernel),
        # so step() has no host-side state and can be replayed inside a CUDA graph
        self._capturable = capturable
        self._step_dev = None
        # fuse_refresh: 2-D weights of `modules` are updated by gdmcf_adamw_refresh, which also rewrites the tensors the
        # contractions derive from them (same update arithmetic; False keeps one flat pass per parameter)
        self._fuse_refresh = fuse_refresh
        self._lazy = {}  # id(param) -> row-sparse parameters updated with exact catch-up (update_rows_lazy)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        self.begin_step()
        self.update(None, grad_scale)
        self.end_step()
        return loss

    # -- the three phases of step(), exposed so that a data-parallel engine can update parameter groups as their
    #    gradient all-reduces complete (engine.StepEngine)
    @torch.no_grad()
    def begin_step(self) -> None:
        self._adopted = []
        if self._capturable:
            if self._step_dev is None:
                dev = next(p for g in self.param_groups for p in g["params"]).device
                self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            K.counter_add(self._step_dev, 1)

    @torch.no_grad()
    def update(self, params=None, grad_scale: float = 1.0, row_coef=None, partition_ctas: int = 0) -> None:
        """AdamW on `params` (default: every parameter with a gradient). row_coef: {id(param): coef [rows]} — the
        parameter's effective gradient is grad + coef[r] * param[r, :] (deferred norm term of the cosine scorer).
        partition_ctas > 0: update only (derived tensors are NOT refreshed — the caller runs kernels.refresh_derived
        afterwards) with gdmcf_adamw_partitioned on that many SMs, so that the pass can overlap SM-limited contractions."""
        row_coef = row_coef or {}
        only = None if params is None else {id(p) for p in params}
        # weights whose derived tensors (bf16 operands, transposes, norms, one-hot tables) are refreshed by the same pass
        specs = {}
        if self._fuse_refresh:
            for mod in self._modules:
                for pid, (kw, names) in mod.refresh_specs().items():
                    specs[pid] = (kw, names, mod)
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None or (only is not None and id(p) not in only):
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                hyper = dict(lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"], weight_decay=group["weight_decay"],
                             step=st["step"], step_dev=self._step_dev if self._capturable else None, grad_scale=grad_scale)
                spec = specs.get(id(p))
                rc = row_coef.get(id(p))
                if partition_ctas > 0:
                    g = p.grad if (p.dim() == 2 and p.grad.stride(1) == 1) or p.grad.is_contiguous() else p.grad.contiguous()
                    K.adamw_partitioned(p.data, g, st["exp_avg"], st["exp_avg_sq"], n_ctas=partition_ctas, **hyper, row_coef=rc)
                elif spec is not None and p.dim() == 2 and p.is_contiguous() and p.grad.stride(1) == 1:
                    K.adamw_refresh(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], **hyper, **spec[0], row_coef=rc)
                    self._adopted.append((spec[2], p, spec[1]))
                else:
                    if rc is not None:
                        p.grad.addcmul_(rc[:, None], p.data)
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    K.adamw_fused(p.data, g, st["exp_avg"], st["exp_avg_sq"], **hyper)

    @torch.no_grad()
    def update_rows(self, p, grad, r0: int, r1: int, grad_scale: float = 1.0, row_coef=None, refresh=None) -> None:
        """AdamW on rows [r0, r1) of the 2-D parameter p only (a rank's shard of a reduce-scattered gradient). `grad` is
        the full-shape gradient view whose rows [r0, r1) hold the reduced values. Derived tensors are NOT refreshed (the
        caller all-gathers the parameter and calls kernels.refresh_derived on the whole matrix) unless `refresh` =
        dict(op_hi=[>= rows, ld] bf16, inv=[>= rows] fp32, rowpart=workspace) is given: then rows [r0, r1) of the bf16
        operand and of the inverse row norms are produced by the same pass (the caller all-gathers THOSE)."""
        assert p.dim() == 2 and p.is_contiguous() and grad.shape == p.shape and grad.stride(1) == 1
        group = next(g for g in self.param_groups if any(q is p for q in g["params"]))
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        st["step"] += 1
        r1 = min(r1, p.shape[0])
        if r1 <= r0:
            return
        b1, b2 = group["betas"]
        derived = {}
        if refresh is not None:
            derived = dict(op=K.Bf16Mat(refresh["op_hi"][r0:r1], None, r1 - r0, p.shape[1]), inv_norm=refresh["inv"][r0:r1],
                           rowpart=refresh["rowpart"])
        K.adamw_refresh(p.data[r0:r1], grad[r0:r1], st["exp_avg"][r0:r1], st["exp_avg_sq"][r0:r1], lr=group["lr"], beta1=b1,
                        beta2=b2, eps=group["eps"], weight_decay=group["weight_decay"], step=st["step"],
                        step_dev=self._step_dev if self._capturable else None, grad_scale=grad_scale,
                        row_coef=row_coef[r0:r1] if row_coef is not None else None, **derived)

    # -- row-sparse parameters (embedding_user: only the batch's rows carry a gradient) ---------------------------------
    def _lazy_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        if "last_step" not in st:
            # rows are up to date with every step taken so far (they were updated densely, or never)
            st["last_step"] = torch.full((p.shape[0],), int(st["step"]), dtype=torch.int32, device=p.device)
            if self._capturable and self._step_dev is not None:
                st["last_step"] += (self._step_dev.to(torch.int32) - int(st["step"]))
        elif st["last_step"].dtype != torch.int32:  # Optimizer.load_state_dict casts state tensors to the parameter's dtype
            st["last_step"] = st["last_step"].to(torch.int32)
        self._lazy[id(p)] = p
        group = next(g for g in self.param_groups if any(q is p for q in g["params"]))
        b1, b2 = group["betas"]
        return st, dict(lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"], weight_decay=group["weight_decay"],
                        step_dev=self._step_dev if self._capturable else None)

    @torch.no_grad()
    def catch_up_rows(self, p, idx, n_sel: int) -> None:
        """Bring rows idx of a lazily updated table up to the last completed step (call BEFORE a forward pass reads them)."""
        if self._capturable and self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=p.device)
        st, hyper = self._lazy_state(p)
        if st["step"] == 0 and not self._capturable:
            return
        K.adamw_rows_lazy(p.data, st["exp_avg"], st["exp_avg_sq"], st["last_step"], idx=idx, n_sel=n_sel, step=max(st["step"], 1),
                          **hyper)

    @torch.no_grad()
    def update_rows_lazy(self, p, idx, grad_rows, n_sel: int, grad_scale: float = 1.0, first: bool = True) -> None:
        """AdamW step on a table whose gradient is non-zero on rows idx only (distinct), grad_rows [n_sel, cols]: the other
        rows are NOT touched now; their zero-gradient updates are replayed exactly when they are next selected, caught up
        or flushed (gdmcf_adamw_rows_lazy). Call between begin_step() and end_step(); `first=False` for further row sets of
        the same step (data parallel: one set per rank)."""
        st, hyper = self._lazy_state(p)
        if first:
            st["step"] += 1
        K.adamw_rows_lazy(p.data, st["exp_avg"], st["exp_avg_sq"], st["last_step"], idx=idx, grad_rows=grad_rows, n_sel=n_sel,
                          step=st["step"], grad_scale=grad_scale, **hyper)

    @torch.no_grad()
    def flush_lazy(self) -> None:
        """Replay the pending zero-gradient steps of every lazily updated table. Required before such a table is read
        outside the training step (evaluation of other users, state_dict, checkpoint)."""
        for p in self._lazy.values():
            st, hyper = self._lazy_state(p)
            if st["step"] > 0 or self._capturable:
                K.adamw_rows_lazy(p.data, st["exp_avg"], st["exp_avg_sq"], st["last_step"], step=max(st["step"], 1), **hyper)

    @torch.no_grad()
    def end_step(self) -> None:
        for m in self._modules:
            m.weights_updated()
        for mod, p, names in self._adopted:
            mod.adopt_refreshed(p, names)
        self._adopted = []
