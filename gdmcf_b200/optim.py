"""Fused AdamW (torch.optim.AdamW semantics, main.py:258,351) on the gdmcf_adamw_fused kernel: one launch per
parameter, p/m/v/grad streamed once (28 B per element). Parameters without a gradient are skipped exactly like
torch (DNNOneHotEmbeddingGCN.out_layers never receives one). `grad_scale` folds the 1/world_size of a
data-parallel gradient sum into the update."""
from __future__ import annotations

import torch

from . import kernels as K


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, modules=(), capturable=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._modules = list(modules)  # engine modules whose bf16 operand caches must be invalidated
        # capturable: the step count lives on the device (one counter shared by all parameters, advanced by a kernel),
        # so step() has no host-side state and can be replayed inside a CUDA graph
        self._capturable = capturable
        self._step_dev = None

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        if self._capturable:
            if self._step_dev is None:
                dev = next(p for g in self.param_groups for p in g["params"]).device
                self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            K.counter_add(self._step_dev, 1)
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                K.adamw_fused(p.data, g, st["exp_avg"], st["exp_avg_sq"], lr=group["lr"], beta1=b1, beta2=b2,
                              eps=group["eps"], weight_decay=group["weight_decay"], step=st["step"],
                              step_dev=self._step_dev if self._capturable else None, grad_scale=grad_scale)
        for m in self._modules:
            m.weights_updated()
        return loss
