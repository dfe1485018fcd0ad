"""Data-parallel plumbing (torch.distributed; NCCL over NVLink on GPUs, gloo in the CPU tests).

The path shards by logical user batch with replicated weights: the only exchanges are one
all-reduce(SUM) of the flat gradient per optimizer step and one all-reduce of the metric sums per evaluation.
The reference has no distributed code at all; a G-rank step equals the reference with gradients averaged over G
consecutive batches before one AdamW step."""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as td


@dataclass
class Dist:
    rank: int = 0
    world_size: int = 1
    local_rank: int = 0
    owns_group: bool = False
    small_group: object = None  # second communicator: small messages do not queue behind a 400 MB all-reduce

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world_size > 1:
            td.all_reduce(t, op=td.ReduceOp.SUM)
        return t

    def all_reduce_gradients(self, model: torch.nn.Module) -> None:
        """Sum gradients across ranks (the 1/world_size is folded into FusedAdamW's grad_scale)."""
        self.all_reduce_tensors([p.grad for p in model.parameters() if p.grad is not None])

    def all_reduce_tensors(self, tensors) -> None:
        """In-place all-reduce(SUM) of a list of tensors: large ones one by one (no 1 GB staging copy of the 400 MB
        embedding gradient), the small ones flattened into one message."""
        if self.world_size == 1:
            return
        small = []
        for g in tensors:
            if not g.is_contiguous():  # [rows, cols] view of a padded [rows, ld] wgrad buffer: reduce the whole buffer
                assert g.dim() == 2 and g.stride(1) == 1
                g = g.as_strided((g.shape[0], g.stride(0)), (g.stride(0), 1))
            if g.numel() * g.element_size() >= (8 << 20):
                td.all_reduce(g, op=td.ReduceOp.SUM)
            else:
                small.append(g)
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            td.all_reduce(flat, op=td.ReduceOp.SUM)
            off = 0
            for g in small:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def all_reduce_async(self, tensors, group=None):
        """Starts the all-reduce(SUM) of `tensors` on NCCL's stream (it waits for the work already queued on the current
        stream). Returns (works, after): call w.wait() for every work, then every function of `after` (they copy the
        flattened small tensors back)."""
        works, after, small = [], [], []
        if self.world_size == 1:
            return works, after
        for g in tensors:
            if not g.is_contiguous():
                assert g.dim() == 2 and g.stride(1) == 1
                g = g.as_strided((g.shape[0], g.stride(0)), (g.stride(0), 1))
            if g.numel() * g.element_size() >= (8 << 20):
                works.append(td.all_reduce(g, op=td.ReduceOp.SUM, group=group, async_op=True))
            else:
                small.append(g)
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            works.append(td.all_reduce(flat, op=td.ReduceOp.SUM, group=group, async_op=True))

            def copy_back():
                off = 0
                for g in small:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
            after.append(copy_back)
        return works, after

    def broadcast_parameters(self, model: torch.nn.Module) -> None:
        if self.world_size == 1:
            return
        for p in model.parameters():
            td.broadcast(p.data, src=0)

    def exchange_factors(self, send_rows, send_small, recv_rows, recv_small, group=None):
        """Exchange of a rank-B gradient's factors (engine.StepEngine: dE = Gs^T hc' summed over ranks, row-sharded):
        send_rows [G * R, B] — block q of its rows goes to rank q (all-to-all) -> recv_rows [G, R, B] (slot r = rank r's
        block of MY rows); send_small [C, B] goes to everyone (all-gather) -> recv_small [G, C, B]. Rank q then forms
        sum_r recv_rows[r] @ recv_small[r]^T = rows [q * R, (q + 1) * R) of the summed product. Returns the async works."""
        G = self.world_size
        assert recv_rows.shape[0] == G and recv_small.shape[0] == G and send_rows.numel() == recv_rows.numel()
        return [td.all_to_all_single(recv_rows.view(-1), send_rows.view(-1), group=group, async_op=True),
                td.all_gather_into_tensor(recv_small.view(-1), send_small.view(-1), group=group, async_op=True)]

    def barrier(self) -> None:
        if self.world_size > 1:
            td.barrier()

    def shutdown(self) -> None:
        if self.owns_group and td.is_initialized():
            td.destroy_process_group()


def init(backend: str | None = None) -> Dist:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). Single process when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return Dist()
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    owns = False
    if not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        opts = None
        if backend == "nccl" and os.environ.get("GDMCF_NCCL_HIGH_PRIORITY", "1") != "0":
            # the all-reduces overlap with large-grid compute kernels: on a high-priority stream NCCL's CTAs take the
            # SM slots that free up first instead of queueing behind the compute kernel's pending CTAs
            opts = td.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
        td.init_process_group(backend=backend, rank=rank, world_size=world, pg_options=opts)
        owns = True
        return Dist(rank, world, local, owns, td.new_group(backend=backend, pg_options=opts))
    return Dist(rank, world, local, owns, td.new_group(backend=backend))
