"""StepEngine — one logical user batch through the whole hot path as CUDA graph replays.

The reference's loop body (main.py:331-351 train step, main.py:288-307 evaluate step) is ~150 kernel launches here;
issued one by one from Python the step is bound by host enqueue time, not by the GPU. The engine captures

    training_losses + backward -> p_sample (reverse loop) -> history mask -> top-k -> Recall/NDCG sums
                               -> FusedAdamW update (+ refresh of the bf16 operands / tables derived from the weights)

once, around *static* device inputs (the batch's CSR rows, ground-truth rows and user ids), and replays it per batch.
Everything that changes from step to step lives on the device: the Philox epoch and AdamW step counters are advanced
by kernels inside the graph, timesteps are drawn by gdmcf_sample_timesteps, bf16 weight operands are refreshed in place.
With world_size > 1 the step is a chain of graph segments with eager NCCL between them: the backward pass hands over
gradient groups as they become final (item table first); the big matrices are reduce-scattered by row blocks, a side
stream runs AdamW on the rank's block and all-gathers the weights while the main stream denoises and ranks, and the
derived tensors are recomputed from the gathered weights.

The training half is `train_step.fused_train_stages` — the kernels and arithmetic of `diffusion.training_losses(...)
["loss"].mean().backward()` without the autograd bookkeeping; `optimizer.update`, `diffusion.rank` and
`metrics_from_device` are the public calls. `graphs=False` runs the same program without capture (parity tests)."""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as td

from . import _lib, evaluate_utils
from . import kernels as K
from .models.gaussian_diffusion import CsrBatch
from .train_step import fused_train_stages


class StepEngine:
    def __init__(self, model, diffusion, optimizer, dist, *, batch_size: int, n_item: int, topk: int, topN: Sequence[int],
                 cap_train_nnz: int, cap_gt_nnz: int, reweight: bool = True, graphs: bool = True, device=None,
                 rank_before_update: bool = True, nccl_sms: int = 0, shard_optimizer: bool = True,
                 shard_min_bytes: int = 64 << 20, train: bool = True, overlap_sms: int = 0,
                 lazy_user_rows: bool = True, rank: bool = True, sampling_steps: int = 0, sampling_noise: bool = False,
                 factor_exchange: bool = True, bf16_gather: bool = True):
        self.model, self.diffusion, self.opt, self.dist = model, diffusion, optimizer, dist
        self.B, self.n_item, self.k, self.topN, self.reweight = batch_size, n_item, topk, list(topN), reweight
        self.use_graphs = graphs
        self.train = train  # False: denoise + rank only (BASELINE.json configs[2]); no gradients, no collectives
        self.do_rank = rank or not train  # False: training step only (main.py's epoch loop, main.py:331-351)
        self.sampling_steps = sampling_steps  # forward-process steps before the reverse loop (--sampling_steps, main.py:288)
        self.sampling_noise = sampling_noise  # --sampling_noise: stochastic reverse steps (gaussian_diffusion.py:745-750)
        self._res = None  # device-resident interaction matrices bound with bind_resident()
        # the item table's norm-term gradient (-E_i * ri^2 * c_i) is applied inside the AdamW pass instead of the wgrad
        # contraction's epilogue (saves a 412 MB read of E per step at the Yelp shape)
        self.defer_item_norm = hasattr(model, "embedding_item")
        # order inside a step: train (forward + backward) -> denoise + rank -> optimizer update (True), or the update
        # before the ranking (False: the batch is ranked with the weights that already include its own gradient)
        self.rank_before_update = rank_before_update
        self.nccl_sms = nccl_sms  # SMs the contractions leave free while all-reduces are in flight (world_size > 1)
        # world_size == 1: SMs (whole TPCs) that run the AdamW pass of the big matrices on a side stream WHILE the
        # tensor-bound denoise + rank phase runs on the others (0: optimizer after the ranking, on the same stream).
        # Measured on B200 (profiles/r2_overlap_sweep.txt): a loss at every split — the exact-rounding AdamW arithmetic
        # needs ~80 instructions per element, so confined to 40-72 SMs the pass is issue-bound at ~43 GB/s per SM
        # (6.07 / 4.95 / 4.30 ms per step at 40 / 56 / 72 SMs against 4.17 serial). Off by default.
        self.overlap_sms = overlap_sms & ~1
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
        self._segments = []          # [(CUDAGraph, communication action after it or None)]
        self._works, self._after, self._result = {}, {}, None
        self._opt_stream = torch.cuda.Stream(device=dev)   # side stream of the sharded optimizer (world_size > 1)
        self._main_stream = None
        self._small_keys = set()     # gradient groups exchanged over dist.small_group
        self.launches_per_step = 0
        # data parallel: the user table's gradient has B non-zero rows per rank -> all-gather (ids, rows), 1.6 MB instead
        # of all-reducing the dense [n_user, d] gradient (218 MB at the Yelp shape)
        G = dist.world_size
        # data parallel: the big matrices' gradients are reduce-scattered by row blocks, each rank runs AdamW on its block
        # only (1/G of the optimizer pass) and the updated blocks are all-gathered; the derived tensors are then refreshed
        # from the gathered weights. Same communication volume as an all-reduce, 1/G of the optimizer's HBM traffic.
        self._shards = {}
        # ... except the item table (half of all gradient bytes): its gradient is the rank-B product Gs^T hc' (B = batch rows),
        # so the ranks exchange the FACTORS instead of the [n_item, 3d] product — every rank receives the row block of each
        # peer's Gs^T that belongs to its shard (all-to-all, n_item/G x B bf16 per peer) plus each peer's hc'^T (all-gather)
        # and forms its block of the SUMMED gradient as one contraction whose K runs over the ranks (one K segment per
        # rank). Same flops as the local gradient contraction it replaces; 24 MB sent per rank and step instead of 360 MB
        # (Yelp shape, 8 ranks). bf16 mode only (the operands of the local contraction are these bf16 factors anyway).
        self.factor_exchange = bool(factor_exchange) and G <= _lib.MAX_SEG and not getattr(model, "_lo", True)
        # ... and its updated rows travel back as the bf16 operand the contractions read (plus the row norms) instead of
        # the fp32 master weights: half the all-gather bytes of the largest matrix, and the owner derives the operand
        # while the updated values are still in registers. The fp32 master of the item table is then current only on
        # the rows a rank owns; flush() all-gathers it (checkpoints, state_dict, evaluation of the weights elsewhere).
        self.bf16_gather = bool(bf16_gather) and not getattr(model, "_lo", True)
        self._stale_masters = False
        self._bf16_gather_active = False
        if G > 1 and shard_optimizer and train:
            # GDMCF_SHARD_MIN_BYTES: size from which a matrix is row-sharded (tests shard toy-sized models through main.py)
            self._setup_shards(int(os.environ.get("GDMCF_SHARD_MIN_BYTES", shard_min_bytes)))
        self.sparse_user_rows = G > 1 and hasattr(model, "embedding_user")
        # The user table's gradient has B non-zero rows per rank. lazy_user_rows: AdamW touches only those rows; the
        # zero-gradient updates every other row would have received (decaying moments still move the weights) are replayed
        # exactly when a row is next used (optim.FusedAdamW.update_rows_lazy; call flush() before reading the table
        # outside step()). Saves a 28 B/element pass over the [n_user, d] table and its dense gradient every step.
        self.lazy_user_rows = bool(lazy_user_rows) and train and hasattr(model, "embedding_user")
        if self.sparse_user_rows:
            d = model.embedding_user.weight.shape[1]
            self._send_idx = torch.zeros(batch_size, **i32)
            self._send_rows = torch.zeros(batch_size, d, dtype=torch.float32, device=dev)
            self._recv_idx = torch.zeros(G, batch_size, **i32)
            self._recv_rows = torch.zeros(G, batch_size, d, dtype=torch.float32, device=dev)

    def _setup_shards(self, min_bytes: int) -> None:
        G, dev = self.dist.world_size, self.dev
        views = {}
        for name, p in self.model.named_parameters():
            if p.dim() != 2 or p.numel() * 4 < min_bytes or name.startswith("embedding_user") or name.startswith("out_layers"):
                continue
            rows, cols = p.shape
            R = (rows + G - 1) // G
            gbuf = torch.zeros(R * G, K.round_up(cols, 4), dtype=torch.float32, device=dev)  # 16 B-aligned rows (TMA store)
            pbuf = torch.zeros(R * G, cols, dtype=torch.float32, device=dev)
            pbuf[:rows].copy_(p.data)
            p.data = pbuf[:rows]  # the parameter now lives in a row-padded buffer: equal blocks for the all-gather
            views[name] = gbuf[:rows, :cols]
            self._shards[name] = dict(p=p, R=R, gbuf=gbuf, pbuf=pbuf, gview=views[name])
            if name == "embedding_item.weight" and self.bf16_gather:
                sh_ = self._shards[name]
                sh_["hi_sh"] = torch.zeros(R * G, K.round_up(cols, 64), dtype=torch.bfloat16, device=dev)
                sh_["inv_sh"] = torch.zeros(R * G, dtype=torch.float32, device=dev)
                sh_["rowpart_sh"] = torch.empty(K.round_up(cols, 64) // 64 * R, dtype=torch.float32, device=dev)  # any split count
            if name == "embedding_item.weight" and self.factor_exchange and self.defer_item_norm:
                Bp = K.round_up(self.B, 64)
                z = lambda *shape: torch.zeros(*shape, dtype=torch.bfloat16, device=dev)  # noqa: E731
                # sendA: Gs^T [G * R, Bp] (rows >= n_item and columns >= B stay zero); sendB: hc'^T [3d, Bp]
                self._shards[name]["fx"] = dict(sendA=z(R * G, Bp), sendB=z(cols, Bp), recvA=z(G, R, Bp), recvB=z(G, cols, Bp))
                self.model._item_factor_send = (self._shards[name]["fx"]["sendA"], self._shards[name]["fx"]["sendB"])
        self.model._grad_views = views
        self.model.weights_updated()

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

    def bind_resident(self, x_dev, gt_dev=None, hist_dev=None, hist2_dev=None) -> None:
        """Resident mode (call before capture): the step reads the batch's rows straight from device-resident interaction
        matrices (data_utils.DeviceInteractions) through the user ids — x_dev: the model input rows (main.py:153-156),
        gt_dev: ground truth of the metrics, hist_dev / hist2_dev: rows masked before top-k (main.py:299; default x_dev).
        Per step only the B user ids change (load_users); nothing is copied or re-indexed."""
        h1 = hist_dev if hist_dev is not None else x_dev
        self._res = dict(x=(x_dev.rowptr, x_dev.col), gt=(gt_dev.rowptr, gt_dev.col) if gt_dev is not None else None,
                         h1=(h1.rowptr, h1.col), h2=(hist2_dev.rowptr, hist2_dev.col) if hist2_dev is not None else None)

    def load_users(self, users) -> None:
        """Resident mode: global user ids int32 [B] (device or pinned host tensor) of the next step."""
        assert self._res is not None and users.numel() == self.B
        self.users.copy_(users, non_blocking=True)

    def load_host(self, users, tr, gt) -> None:
        """Pinned host tensors: users int32 [B]; tr / gt = (rowptr int32 [B+1] starting at 0, col int32 [nnz])."""
        self.users.copy_(users, non_blocking=True)
        for (rp_h, cl_h), rp, cl in ((tr, self.tr_rowptr, self.tr_col), (gt, self.gt_rowptr, self.gt_col)):
            assert cl_h.numel() <= cl.numel(), "batch has more interactions than the engine's capacity"
            rp.copy_(rp_h, non_blocking=True)
            cl[: cl_h.numel()].copy_(cl_h, non_blocking=True)

    # -- the step ----------------------------------------------------------------------------------
    def _batch(self) -> CsrBatch:
        if self._res is not None:
            return CsrBatch(self._res["x"][0], self._res["x"][1], self.users, self.n_item)
        return CsrBatch(self.tr_rowptr, self.tr_col, self.local_ids, self.n_item)

    def _rank_and_metrics(self):
        """p_sample -> history mask -> top-k -> metric sums (main.py:288-307) on the loaded batch."""
        model, diff = self.model, self.diffusion
        model.eval()
        if not self.train and hasattr(model, "invalidate_time_tables"):
            # an inference-only program may run after optimizer steps of another program (main.py: evaluation after an
            # epoch): the bias tables are rebuilt inside the step; the bf16 operands are kept current by the optimizer pass
            model.invalidate_time_tables()
        batch = self._batch()
        if self._res is not None:
            idx = diff.rank(model, batch, self.k, hist=self._res["h1"], hist2=self._res["h2"], steps=self.sampling_steps,
                            index=self.users, sampling_noise=self.sampling_noise)
            gt = self._res["gt"]
        else:
            idx = diff.rank(model, batch, self.k, hist=(self.tr_rowptr, self.tr_col), steps=self.sampling_steps, index=self.users,
                            sampling_noise=self.sampling_noise)
            gt = (self.gt_rowptr, self.gt_col)
        if gt is None:
            return idx, None
        return idx, evaluate_utils.metrics_from_device(idx, batch.users, gt[0], gt[1], self.topN)

    def _program(self):
        """The step as a generator. Between two yields everything is device work on the current stream (one CUDA graph
        segment when capturing); a yield is a communication point: ("reduce", key, tensors) starts the asynchronous
        all-reduce of a finished gradient group, ("wait", key) makes the stream wait for it. With one rank there are no
        yields and the whole step is a single graph."""
        model, diff, opt, G = self.model, self.diffusion, self.opt, self.dist.world_size
        if not self.train:
            idx, sums = self._rank_and_metrics()
            self._result = (torch.zeros((), dtype=torch.float64, device=self.dev), idx, sums)
            return
        if (G > 1 or self.overlap_sms > 0) and self.do_rank and getattr(model, "can_project", lambda: False)():
            # The ranking phase runs next to the optimizer stream, which rewrites the fp32 master weights in place (row block
            # update + all-gather). Everything the ranking reads must therefore be derived BEFORE the fork: the projection
            # operand P = W1 diag(ri) E is the one tensor the reverse loop would otherwise build from fp32 W1 on first use.
            with torch.no_grad():
                model._projection_operand()
        model.train()
        opt.zero_grad(set_to_none=True)
        params = dict(model.named_parameters())
        lazy = self.lazy_user_rows
        if lazy:  # the forward pass reads these users' rows: bring them up to the last completed optimizer step first
            opt.catch_up_rows(params["embedding_user.weight"], self.users, self.B)
        stages = fused_train_stages(diff, model, self._batch(), self.reweight, index=self.users,
                                    defer_item_norm=self.defer_item_norm, sparse_user_grad=lazy)
        _, loss = next(stages)
        groups, row_coef = [], {}
        user_gi = None
        for si, (_, grads) in enumerate(stages):
            names = list(grads)
            has_user = "embedding_user.weight" in grads or (lazy and si == 1)
            if has_user:
                user_gi = len(groups)
            for n in names:
                sh = self._shards.get(n)
                if grads[n] is None:  # exchanged as factors (see __init__): this rank's block is formed in finish_group
                    assert sh is not None and "fx" in sh
                    params[n].grad = None
                    continue
                if sh is not None and grads[n].data_ptr() != sh["gview"].data_ptr():
                    sh["gview"].copy_(grads[n])  # gradient not produced in place (no _grad_buffer hook for this weight)
                    grads[n] = sh["gview"]
                params[n].grad = grads[n]
            groups.append([params[n] for n in names])
            if "embedding_item.weight" in grads and getattr(model, "_item_grad_rowcoef", None) is not None:
                row_coef[id(params["embedding_item.weight"])] = model._item_grad_rowcoef
            if G > 1:
                dense = [grads[n] for n in names
                         if not (self.sparse_user_rows and n == "embedding_user.weight") and n not in self._shards]
                fx = [n for n in names if grads[n] is None]
                dense += [row_coef[id(params[n])] for n in names if id(params[n]) in row_coef]  # summed like the gradient
                rs = [n for n in names if n in self._shards and n not in fx]
                if not rs and not fx and sum(t.numel() * t.element_size() for t in dense) < (64 << 20):
                    self._small_keys.add(len(groups) - 1)
                dense = (dense, rs, fx)
                if self.sparse_user_rows and has_user:
                    # only B rows of the user table carry a gradient: exchange (ids, rows) instead of the dense table
                    idx, rows = model._user_grad_rows
                    self._send_idx.copy_(idx)
                    self._send_rows.copy_(rows)
                    yield ("gather_rows", len(groups) - 1, dense)
                else:
                    yield ("reduce", len(groups) - 1, dense)
                if self.nccl_sms > 0 and len(groups) == 1:
                    # all-reduces are in flight from here to the last wait: the contractions leave SMs to NCCL
                    K.gemm_set_sm_limit(max(2, _lib.load().gdmcf_num_sms() - self.nccl_sms))

        def rank_and_metrics():
            return self._rank_and_metrics() if self.do_rank else (None, None)

        order = sorted(range(len(groups)), key=lambda gi: (0 if gi in self._small_keys else 1, gi))
        by_param = {id(sh["p"]): (n, sh) for n, sh in self._shards.items()}
        gathers, refreshed = [], []
        # parameters whose row blocks are all-gathered as bf16 operand rows (see __init__): needs the operand, its transpose
        # and the row norms to exist (they do once the model has run), bf16 mode
        use16 = {}
        if G > 1 and self.bf16_gather:
            for n_, sh_ in self._shards.items():
                spec_ = model.refresh_specs().get(id(sh_["p"])) if "hi_sh" in sh_ else None
                if spec_ is not None and {"op", "op_t", "inv_norm"} <= set(spec_[0]) and spec_[0]["op"].lo is None \
                        and spec_[0]["op"].hi.shape[1] == sh_["hi_sh"].shape[1]:
                    use16[id(sh_["p"])] = spec_
        self._bf16_gather_active = bool(use16)

        def finish_group(gi, sharded_rows: bool, replicated: bool):
            plist = groups[gi]
            if replicated and lazy and gi == user_gi:
                pU = params["embedding_user.weight"]
                if G == 1:
                    idx_, rows_ = model._user_grad_rows
                    opt.update_rows_lazy(pU, idx_, rows_, self.B)
                else:  # one row set per rank (distinct users: the ranks hold different batches), identical on every rank
                    for r in range(G):
                        opt.update_rows_lazy(pU, self._recv_idx[r], self._recv_rows[r], self.B, grad_scale=1.0 / G, first=r == 0)
            if replicated:
                if not lazy and self.sparse_user_rows and any(p is params.get("embedding_user.weight") for p in plist):
                    gU = params["embedding_user.weight"].grad
                    d = gU.shape[1]
                    for r in range(G):
                        if r != self.dist.rank:
                            K.scatter_rows_add(self._recv_rows[r], self._recv_idx[r], gU, self.B, d)
                opt.update([p for p in plist if id(p) not in by_param], grad_scale=1.0 / G, row_coef=row_coef)
            mine = [by_param[id(p)] for p in plist if id(p) in by_param] if sharded_rows else []
            for n, sh in mine:  # this rank's row block of the reduce-scattered gradient
                r0 = self.dist.rank * sh["R"]
                if "fx" in sh and sh["p"].grad is None:
                    # factor exchange: block = sum over ranks of Gs_r^T[block rows] hc'_r, K = (rank, batch row)
                    fx_, cols_ = sh["fx"], sh["p"].shape[1]
                    K.gemm([fx_["recvA"][r_] for r_ in range(G)], [fx_["recvB"][r_] for r_ in range(G)], sh["R"], cols_,
                           [self.B] * G, out_f32=sh["gbuf"][r0:r0 + sh["R"], :cols_], splits=1)
                    # (splits=1: no split-K workspace — this runs on the optimizer stream next to the ranking contractions)
                blk = None
                if id(sh["p"]) in use16:  # derive this block's bf16 operand rows + row norms into the shadow buffers
                    blk = dict(op_hi=sh["hi_sh"], inv=sh["inv_sh"], rowpart=sh["rowpart_sh"])
                opt.update_rows(sh["p"], sh["gview"], r0, r0 + sh["R"], grad_scale=1.0 / G, row_coef=row_coef.get(id(sh["p"])),
                                refresh=blk)
            return mine

        side_opt = bool(self._shards) and self.rank_before_update
        if G == 1 and self.rank_before_update and self.overlap_sms > 0:
            # One rank: fork inside the step (one CUDA graph with two branches). Side stream: AdamW of the big matrices on
            # `overlap_sms` SMs, fp32 masters / moments only. Main stream: denoise + rank on the remaining SMs — it reads
            # only the bf16 operands and tables derived from the PRE-update weights, plus the user rows staged before the
            # fork — then AdamW of the small tensors. After the join one pass re-derives the operands from the new weights.
            opt.begin_step()
            n_sms = _lib.load().gdmcf_num_sms()
            specs = model.refresh_specs()
            every = [p for plist in groups for p in plist]
            side_params = [p for p in every if p.dim() == 2 and p.numel() >= (1 << 20)]
            side_ids = {id(p) for p in side_params}
            if hasattr(model, "stage_user_rows"):
                model.stage_user_rows(self.B, self.users)
            main, side = torch.cuda.current_stream(self.dev), self._opt_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                opt.update(side_params, row_coef=row_coef, partition_ctas=self.overlap_sms)
            K.gemm_set_sm_limit(max(2, n_sms - self.overlap_sms))
            try:
                idx, sums = rank_and_metrics()
            finally:
                K.gemm_set_sm_limit(0)
            opt.update([p for p in every if id(p) not in side_ids], row_coef=row_coef)
            if lazy:
                opt.update_rows_lazy(params["embedding_user.weight"], *model._user_grad_rows, self.B)
            main.wait_stream(side)
            for p_ in side_params:
                spec = specs.get(id(p_))
                if spec is not None:
                    K.refresh_derived(p_.data, **spec[0])
                    refreshed.append((p_, spec[1]))
            opt.end_step()
            for p_, names_ in refreshed:
                model.adopt_refreshed(p_, names_)
            self._result = (loss, idx, sums)
            return
        if side_opt:
            # Sharded optimizer on a side stream: as each reduce-scatter lands, this rank's row block is updated and the
            # all-gather of the weights starts, all concurrently with the denoise + rank phase on the main stream (which
            # reads only the bf16 operands / tables derived from the OLD weights; they are refreshed after the join).
            opt.begin_step()
            big = [gi for gi in order if any(id(p) in by_param for p in groups[gi])]
            yield ("to_opt", None, None)
            for gi in big:
                yield ("wait", gi, None)
                mine = finish_group(gi, True, False)
                yield ("gather", ("ag", gi), [(n, id(sh_["p"]) in use16) for n, sh_ in mine])
                gathers.append((("ag", gi), mine))
            yield ("to_main", None, None)
            idx, sums = rank_and_metrics()
            for gi in order:
                if gi not in big:
                    yield ("wait", gi, None)
                    finish_group(gi, False, True)
            yield ("join_opt", None, None)
            for gi in big:  # the small tensors that travelled with a big group (biases, time-embedding layer)
                finish_group(gi, False, True)
        else:
            if self.rank_before_update:
                # denoise + rank this batch with the weights the training step just used; the optimizer update (and, with
                # several ranks, the wait for the gradient all-reduces) comes after, so the collectives hide behind ~1.3 ms
                # of inference work. One train step and one rank step per batch either way.
                idx, sums = rank_and_metrics()
            opt.begin_step()
            # update order: groups whose exchange went over the small-message communicator first (they are complete long
            # before the big all-reduces), then the others in the order their all-reduces were issued
            for gi in order:
                if G > 1:
                    yield ("wait", gi, None)
                mine = finish_group(gi, True, True)
                if mine:
                    yield ("gather", ("ag", gi), [(n, id(sh_["p"]) in use16) for n, sh_ in mine])
                    gathers.append((("ag", gi), mine))
        specs = model.refresh_specs() if gathers else {}
        for key, mine in gathers:
            yield ("wait", key, None)
            for n, sh in mine:
                spec = specs.get(id(sh["p"]))
                if id(sh["p"]) in use16:
                    # the gathered rows ARE the operand: copy into the live operand (the ranking phase was still reading
                    # it while the all-gather ran), transpose for the dgrad operand, adopt the gathered row norms
                    kw_, names_ = use16[id(sh["p"])]
                    rows_, cols_ = sh["p"].shape
                    kw_["op"].hi[:rows_].copy_(sh["hi_sh"][:rows_])
                    K.transpose_bf16(kw_["op"].hi, rows_, cols_, kw_["op_t"].hi)
                    kw_["inv_norm"].copy_(sh["inv_sh"][:rows_])
                    refreshed.append((sh["p"], names_))
                elif spec is not None:  # derived tensors from the gathered weights (same producer as the fused refresh)
                    K.refresh_derived(sh["p"].data, **spec[0])
                    refreshed.append((sh["p"], spec[1]))
        opt.end_step()
        for p_, names_ in refreshed:
            model.adopt_refreshed(p_, names_)
        if G > 1 and self.nccl_sms > 0:
            K.gemm_set_sm_limit(0)
        if not self.rank_before_update:
            idx, sums = rank_and_metrics()
        self._result = (loss, idx, sums)

    # -- communication actions (eager NCCL between graph segments) -----------------------------------
    def _comm(self, action, key, payload) -> None:
        G, rank = self.dist.world_size, self.dist.rank
        if action == "to_opt":  # fork: the optimizer stream continues from here, the main stream stays free
            self._opt_stream.wait_stream(self._main_stream)
            torch.cuda.set_stream(self._opt_stream)
            return
        if action == "to_main":
            torch.cuda.set_stream(self._main_stream)
            return
        if action == "join_opt":
            self._main_stream.wait_stream(self._opt_stream)
            return
        if action == "wait":
            for w in self._works.pop(key, []):
                w.wait()
            for fn in self._after.pop(key, []):
                fn()
            return
        if action == "gather":  # all-gather of the updated row blocks, in place in the parameter's padded buffer
            works = []
            for n, as_bf16 in payload:
                if as_bf16:
                    for buf in (self._shards[n]["hi_sh"], self._shards[n]["inv_sh"]):
                        works.append(td.all_gather_into_tensor(buf.view(-1), buf.view(G, -1)[rank], async_op=True))
                    continue
                pbuf = self._shards[n]["pbuf"]
                works.append(td.all_gather_into_tensor(pbuf.view(-1), pbuf.view(G, -1)[rank], async_op=True))
            self._works[key], self._after[key] = works, []
            return
        tensors, rs, fx = payload
        group = self.dist.small_group if key in self._small_keys else None
        works, after = self.dist.all_reduce_async(tensors, group=group)
        for n in rs:  # reduce-scatter by row blocks, in place: block `rank` of the buffer receives the sum
            gbuf = self._shards[n]["gbuf"]
            works.append(td.reduce_scatter_tensor(gbuf.view(G, -1)[rank], gbuf.view(-1), op=td.ReduceOp.SUM, async_op=True))
        for n in fx:  # factors of a rank-B gradient: row blocks of Gs^T to their owners, hc'^T to everyone
            b = self._shards[n]["fx"]
            works += self.dist.exchange_factors(b["sendA"], b["sendB"], b["recvA"], b["recvB"])
        if action == "gather_rows":
            works.append(td.all_gather_into_tensor(self._recv_idx.view(-1), self._send_idx, group=group, async_op=True))
            works.append(td.all_gather_into_tensor(self._recv_rows.view(-1), self._send_rows.view(-1), group=group, async_op=True))
        self._works[key], self._after[key] = works, after

    def flush(self) -> None:
        """Bring lazily updated tables (lazy_user_rows) up to the current optimizer step. Call before the parameters are
        read outside step(): evaluation of other users, state_dict(), checkpoints, comparisons."""
        if self.train:
            if self._stale_masters:
                # bf16_gather: the fp32 master rows of the other ranks' blocks were not exchanged during the steps
                G, rank = self.dist.world_size, self.dist.rank
                for sh in self._shards.values():
                    if "hi_sh" in sh:
                        td.all_gather_into_tensor(sh["pbuf"].view(-1), sh["pbuf"].view(G, -1)[rank])
                self._stale_masters = False
            self.opt.flush_lazy()
            self.model.weights_updated()
            if hasattr(self.model, "refresh_inference_operands"):
                self.model.refresh_inference_operands()

    def gather_optimizer_state(self) -> None:
        """Row-sharded optimizer (world_size > 1): every rank holds the AdamW moments of its own row block only. Collective:
        all-gathers the blocks so that optimizer.state_dict() is complete on every rank (checkpoints). No-op with one rank."""
        G, rank = self.dist.world_size, self.dist.rank
        if G == 1 or not self._shards or self.opt is None:
            return
        for sh in self._shards.values():
            st = self.opt.state.get(sh["p"])
            if not st:
                continue
            rows, cols = sh["p"].shape
            R = sh["R"]
            r0, r1 = rank * R, min((rank + 1) * R, rows)
            for key in ("exp_avg", "exp_avg_sq"):
                t = st[key]
                blk = torch.zeros(R, cols, dtype=t.dtype, device=t.device)
                if r1 > r0:
                    blk[:r1 - r0].copy_(t[r0:r1])
                full = torch.empty(G * R, cols, dtype=t.dtype, device=t.device)
                td.all_gather_into_tensor(full.view(-1), blk.view(-1))
                t.copy_(full[:rows])
                del blk, full

    def _eager_step(self):
        self._main_stream = torch.cuda.current_stream(self.dev)
        try:
            for action, key, tensors in self._program():
                self._comm(action, key, tensors)
        finally:
            torch.cuda.set_stream(self._main_stream)
        self._stale_masters = self._stale_masters or self._bf16_gather_active
        return self._result

    # -- state snapshot around the warm-up steps -------------------------------------------------------
    def _snapshot(self):
        opt, diff = self.opt, self.diffusion
        snap = dict(params=[p.detach().clone() for p in self.model.parameters()],
                    lt=(diff.Lt_history.clone(), diff.Lt_count.clone()),
                    epoch=None if diff._epoch is None else diff._epoch.clone(),
                    rng_calls=getattr(self.model, "_rng_calls", 0))
        if opt is not None:
            snap["step_dev"] = None if opt._step_dev is None else opt._step_dev.clone()
            snap["state"] = {id(p): {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()}
                             for g in opt.param_groups for p in g["params"] if opt.state.get(p)}
        return snap

    def _restore(self, snap) -> None:
        opt, diff = self.opt, self.diffusion
        with torch.no_grad():
            for p, v in zip(self.model.parameters(), snap["params"]):
                p.copy_(v)
            diff.Lt_history.copy_(snap["lt"][0]); diff.Lt_count.copy_(snap["lt"][1])
            if diff._epoch is not None:
                diff._epoch.zero_() if snap["epoch"] is None else diff._epoch.copy_(snap["epoch"])
            if hasattr(self.model, "_rng_calls"):
                self.model._rng_calls = snap["rng_calls"]
            if opt is not None:
                if opt._step_dev is not None:
                    opt._step_dev.zero_() if snap["step_dev"] is None else opt._step_dev.copy_(snap["step_dev"])
                for g in opt.param_groups:
                    for p in g["params"]:
                        st, old = opt.state.get(p), snap["state"].get(id(p))
                        if not st:
                            continue
                        for k, v in st.items():  # tensors keep their storage (the captured graph holds their addresses)
                            if torch.is_tensor(v):
                                v.zero_() if old is None else v.copy_(old[k])
                            else:
                                st[k] = 0 if old is None else old[k]
        # the captured step reads the bf16 operands / tables derived from the weights at fixed addresses and only rewrites
        # them in its own optimizer pass: re-derive them in place from the restored weights
        self._stale_masters = False  # every row of every master was just restored from the (complete) snapshot
        self.model.weights_updated()
        by_id = {id(p): p for p in self.model.parameters()}
        for pid, (kw, names) in self.model.refresh_specs().items():
            K.refresh_derived(by_id[pid].data, **kw)
            self.model.adopt_refreshed(by_id[pid], names)

    def capture(self, warmup: int = 3, preserve_state: bool = False) -> None:
        """Eager warm-up on whatever the static inputs hold (call load_* first), then capture. The warm-up allocates the
        persistent operand buffers and optimizer state outside the graph pool and fills Lt_history; the first capture
        needs at least one warm-up step (a re-capture of a warmed engine may pass 0). preserve_state: weights, optimizer
        state, Lt_history and RNG counters are put back (in place) after the warm-up, so that the warm-up steps are not
        part of the training run (main.py); the benchmark lets them count as training steps."""
        if self._stale_masters:
            self.flush()
        snap = self._snapshot() if preserve_state and warmup > 0 else None
        try:
            self._capture(warmup)
        finally:
            if snap is not None:
                self._restore(snap)

    def _capture(self, warmup: int) -> None:
        lib = _lib.load()
        if hasattr(self.model, "build_derived"):
            with torch.no_grad():
                self.model.build_derived()  # the captured optimizer pass refreshes what exists now (see build_derived)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        if not self.use_graphs:
            return
        if self.train:
            assert getattr(self.opt, "_capturable", False), "graph capture needs FusedAdamW(..., capturable=True)"
            self.opt.zero_grad(set_to_none=True)
        n0 = lib.gdmcf_launch_count()
        self._segments = []
        pool = None

        def begin():
            nonlocal pool
            g = torch.cuda.CUDAGraph()
            ctx = torch.cuda.graph(g, pool=pool, stream=side)
            ctx.__enter__()
            return g, ctx

        on_opt = False  # which stream the segment being captured will be replayed on
        g, ctx = begin()
        try:
            for action in self._program():
                ctx.__exit__(None, None, None)
                pool = pool or g.pool()
                self._segments.append((g, on_opt, action))
                if action[0] == "to_opt":
                    on_opt = True
                elif action[0] == "to_main":
                    on_opt = False
                g, ctx = begin()
        except BaseException:
            ctx.__exit__(None, None, None)
            raise
        ctx.__exit__(None, None, None)
        self._segments.append((g, on_opt, None))
        self.launches_per_step = int(lib.gdmcf_launch_count() - n0)
        self.model.weights_updated()  # capture executed nothing: cached operands follow the replays from here on
        torch.cuda.synchronize(self.dev)

    def step(self):
        """Run one step on the loaded inputs. Returns (loss scalar f64, top-k indices int32 [B, k], metric sums f64
        [len(topN), 4]) — device tensors that the next step overwrites."""
        if not self._segments:
            return self._eager_step()
        self._main_stream = torch.cuda.current_stream(self.dev)
        try:
            for g, on_opt, action in self._segments:
                # the stream is set by the fork / join actions themselves; a segment replays on the stream it follows
                g.replay()
                if action is not None:
                    self._comm(*action)
        finally:
            torch.cuda.set_stream(self._main_stream)
        self.model.weights_updated()  # keeps the eager API coherent: its cached operands are stale after a replay
        self._stale_masters = self._stale_masters or self._bf16_gather_active
        return self._result
