"""StepEngine — one logical user batch through the whole hot path as CUDA graph replays.

The reference's loop body (main.py:331-351 train step, main.py:288-307 evaluate step) is ~150 kernel launches here;
issued one by one from Python the step is bound by host enqueue time, not by the GPU. The engine captures

    training_losses -> loss.mean().backward() -> [gradient all-reduce] -> FusedAdamW.step
                    -> p_sample (reverse loop) -> history mask -> top-k -> Recall/NDCG sums

once, around *static* device inputs (the batch's CSR rows, ground-truth rows and user ids), and replays it per batch.
Everything that changes from step to step lives on the device: the Philox epoch and AdamW step counters are advanced
by kernels inside the graph, timesteps are drawn by gdmcf_sample_timesteps, bf16 weight operands are refreshed in place.
With world_size > 1 the step is two graphs with the NCCL all-reduce of the gradient buffers between them.

The public calls (`diffusion.training_losses`, `loss.backward()`, `optimizer.step()`, `diffusion.rank`) are exactly the
ones a user of the eager API makes; `graphs=False` runs the same code without capture (used by parity tests)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, evaluate_utils
from .models.gaussian_diffusion import CsrBatch


class StepEngine:
    def __init__(self, model, diffusion, optimizer, dist, *, batch_size: int, n_item: int, topk: int, topN: Sequence[int],
                 cap_train_nnz: int, cap_gt_nnz: int, reweight: bool = True, graphs: bool = True, device=None):
        self.model, self.diffusion, self.opt, self.dist = model, diffusion, optimizer, dist
        self.B, self.n_item, self.k, self.topN, self.reweight = batch_size, n_item, topk, list(topN), reweight
        self.use_graphs = graphs
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        self.dev = dev
        i32 = dict(dtype=torch.int32, device=dev)
        # static inputs of the captured step
        self.users = torch.zeros(batch_size, **i32)                 # global user ids (embedding_user rows)
        self.tr_rowptr = torch.zeros(batch_size + 1, **i32)         # batch-local CSR of the training rows: x_start and
        self.tr_col = torch.zeros(max(cap_train_nnz, 1), **i32)     # the history mask of the ranking step
        self.gt_rowptr = torch.zeros(batch_size + 1, **i32)         # batch-local CSR of the ground-truth rows
        self.gt_col = torch.zeros(max(cap_gt_nnz, 1), **i32)
        self.local_ids = torch.arange(batch_size, **i32)
        self._g1: Optional[torch.cuda.CUDAGraph] = None
        self._g2: Optional[torch.cuda.CUDAGraph] = None
        self._grads = None
        self._out = None
        self.launches_per_step = 0

    # -- inputs ------------------------------------------------------------------------------------
    def load_resident(self, train_dev, gt_dev, lo: int, hi: int) -> None:
        """Users [lo, hi) from device-resident interaction matrices (data_utils.DeviceInteractions): device-to-device
        slices of rowptr / col into the static buffers."""
        assert hi - lo == self.B
        self.users.copy_(torch.arange(lo, hi, dtype=torch.int32, device=self.dev))
        for src, rp, cl in ((train_dev, self.tr_rowptr, self.tr_col), (gt_dev, self.gt_rowptr, self.gt_col)):
            b, e = int(src.rowptr_host[lo]), int(src.rowptr_host[hi])
            assert e - b <= cl.numel(), "batch has more interactions than the engine's capacity"
            torch.sub(src.rowptr[lo:hi + 1], b, out=rp)
            if e > b:
                cl[: e - b].copy_(src.col[b:e])

    def load_host(self, users, tr, gt) -> None:
        """Pinned host tensors: users int32 [B]; tr / gt = (rowptr int32 [B+1] starting at 0, col int32 [nnz])."""
        self.users.copy_(users, non_blocking=True)
        for (rp_h, cl_h), rp, cl in ((tr, self.tr_rowptr, self.tr_col), (gt, self.gt_rowptr, self.gt_col)):
            assert cl_h.numel() <= cl.numel(), "batch has more interactions than the engine's capacity"
            rp.copy_(rp_h, non_blocking=True)
            cl[: cl_h.numel()].copy_(cl_h, non_blocking=True)

    # -- the step ----------------------------------------------------------------------------------
    def _batch(self) -> CsrBatch:
        return CsrBatch(self.tr_rowptr, self.tr_col, self.local_ids, self.n_item)

    def _train_part(self):
        self.model.train()
        self.opt.zero_grad(set_to_none=True)
        losses = self.diffusion.training_losses(self.model, self._batch(), self.reweight, index=self.users)
        loss = losses["loss"].mean()
        loss.backward()
        return loss.detach()

    def _update_part(self):
        self.opt.step(grad_scale=1.0 / self.dist.world_size)
        self.model.eval()
        batch = self._batch()
        idx = self.diffusion.rank(self.model, batch, self.k, hist=(self.tr_rowptr, self.tr_col), index=self.users)
        sums = evaluate_utils.metrics_from_device(idx, batch.users, self.gt_rowptr, self.gt_col, self.topN)
        return idx, sums

    def _eager_step(self):
        loss = self._train_part()
        if self.dist.world_size > 1:
            self.dist.all_reduce_tensors([p.grad for p in self.model.parameters() if p.grad is not None])
        idx, sums = self._update_part()
        return loss, idx, sums

    def capture(self, warmup: int = 3) -> None:
        """Eager warm-up on whatever the static inputs hold (call load_* first), then capture. The warm-up allocates the
        persistent operand buffers and optimizer state outside the graph pool and fills Lt_history."""
        lib = _lib.load()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager_step()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        if not self.use_graphs:
            return
        assert getattr(self.opt, "_capturable", False), "graph capture needs FusedAdamW(..., capturable=True)"
        self.opt.zero_grad(set_to_none=True)
        n0 = lib.gdmcf_launch_count()
        self._g1 = torch.cuda.CUDAGraph()
        if self.dist.world_size == 1:
            with torch.cuda.graph(self._g1):
                loss = self._train_part()
                idx, sums = self._update_part()
        else:
            with torch.cuda.graph(self._g1):
                loss = self._train_part()
            self._grads = [p.grad for p in self.model.parameters() if p.grad is not None]
            self._g1.replay()  # gradients of the current static batch (capture itself executes nothing)
            self.dist.all_reduce_tensors(self._grads)
            self._g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g2, pool=self._g1.pool()):
                idx, sums = self._update_part()
            self._g2.replay()
        self.launches_per_step = int(lib.gdmcf_launch_count() - n0)
        self._out = (loss, idx, sums)
        self.model.weights_updated()  # operands were refreshed in place by the replays; eager users must rebuild
        torch.cuda.synchronize(self.dev)

    def step(self):
        """Run one step on the loaded inputs. Returns (loss scalar f64, top-k indices int32 [B, k], metric sums f64
        [len(topN), 4]) — device tensors that the next step overwrites."""
        if self._g1 is None:
            return self._eager_step()
        self._g1.replay()
        if self._g2 is not None:
            self.dist.all_reduce_tensors(self._grads)
            self._g2.replay()
        self.model.weights_updated()  # keeps the eager API coherent: its cached operands are stale after a replay
        return self._out
