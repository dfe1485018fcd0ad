"""Train a diffusion model for recommendation — mirror of the reference's main.py:108-384 (same flow, prints and
model-selection rule), rebuilt CSR-native on the B200 engine:

  * no dense n_user x n_item host tensors (main.py:143-156, D8): interactions live on the GPU as CSR;
  * the train-loop body of main.py:331-351 actually runs (the reference's stray `continue` at :328 is D2);
  * evaluate() (main.py:267-310) is fused: p_sample -> history mask -> top-K -> metric sums, all on the device;
  * `--n_user` replaces the hard-coded `n_user = 3000` (D3); 0 means all users;
  * data-parallel over logical user batches when launched under torchrun (gradient all-reduce + metric all-reduce
    over NCCL; everything else is rank-local), see gdmcf_b200/dist_utils.py;
  * by default the loop body and evaluate() are engine.StepEngine programs (the training step / the denoise + rank step
    captured once as CUDA graphs and replayed per batch, overlapped collectives under torchrun); `--eager` issues the
    same kernels call by call through the reference-shaped API (training_losses -> backward -> optimizer.step).
"""
from __future__ import annotations

import os
import sys
import time
from datetime import datetime

import numpy as np
import torch

from . import data_utils, dist_utils, evaluate_utils
from .models import gaussian_diffusion as gd
from .engine import StepEngine
from .models.DNN import DNN, DNNOneHotEmbeddingGCN
from .optim import FusedAdamW
from .parse_args_util import parse_args

SYNTHETIC_SHAPES = {"yelp": (54574, 34395, 1402736, 0), "amazon": (108822, 94949, 3146256, 1),
                    "scaled": (1000000, 200000, 50000000, 2)}


def load_interactions(args):
    if args.synthetic:
        if args.synthetic in SYNTHETIC_SHAPES:
            U, I, P, seed = SYNTHETIC_SHAPES[args.synthetic]
        else:
            U, I, P = (int(x) for x in args.synthetic.split(","))
            seed = 0
        tr, va, te = data_utils.synthetic_interactions(U, I, P, seed)
        import scipy.sparse as sp
        n_user, n_item = int(tr[:, 0].max()) + 1, int(tr[:, 1].max()) + 1
        mk = lambda p: sp.csr_matrix((np.ones(len(p)), (p[:, 0], p[:, 1])), dtype="float64", shape=(n_user, n_item))  # noqa: E731
        print(f'user num: {n_user}')
        print(f'item num: {n_item}')
        return mk(tr), mk(va), mk(te), n_user, n_item
    return data_utils.data_load(args.data_path + 'train_list.npy', args.data_path + 'valid_list.npy',
                                args.data_path + 'test_list.npy')


def build(args, n_user_rows, n_item, device):
    if args.mean_type == 'x0':
        mean_type = gd.ModelMeanType.START_X
    elif args.mean_type == 'eps':
        mean_type = gd.ModelMeanType.EPSILON
    else:
        raise ValueError("Unimplemented mean type %s" % args.mean_type)
    CatOneHot = args.OneHotMatrix == 2
    diffusion = gd.GaussianDiffusionDiscrete(mean_type, args.noise_schedule, args.noise_scale, args.noise_min,
                                             args.noise_max, args.steps, device, discrete=args.discrete,
                                             CatOneHot=CatOneHot, args=args).to(device)
    out_dims = args.dims + [n_item]
    in_dims = out_dims[::-1]
    print('in_dims:{} out_dims:{}'.format(in_dims, out_dims))
    print('backbone:', args.backbone)
    if args.backbone == 'DNN':
        if CatOneHot:
            raise ValueError("backbone DNN needs OneHotMatrix != 2")
        model = DNN(in_dims, out_dims, args.emb_size, time_type="cat", norm=args.norm, precision=args.precision).to(device)
    elif args.backbone == 'DNNOneHotEmbeddingGCN':
        diffusion.indexIn = True
        model = DNNOneHotEmbeddingGCN(in_dims, out_dims, args.emb_size, time_type="cat", norm=args.norm, item_num=n_item,
                                      user_num=n_user_rows, args=args, precision=args.precision).to(device)
    else:
        print('not implemented!')  # the other --backbone values of main.py:212-256 are outside the hot path
        sys.exit(1)
    return diffusion, model


def evaluate(diffusion, model, train_dev, gt_dev, hist_devs, n_user, batch_size, topN, sampling_steps, dist, drop_last=True,
             sampling_noise=False):
    """main.py:267-310: rank the first floor(n_user / batch_size) * batch_size users (drop_last=True, :156); every user
    with drop_last=False (the --tst_w_val loader, main.py:174)."""
    model.eval()
    k = topN[-1]
    n_batches = n_user // batch_size if drop_last else -(-n_user // batch_size)
    sums = torch.zeros(len(topN), 4, dtype=torch.float64, device=train_dev.device)
    for b in range(dist.rank, n_batches, dist.world_size):
        users = torch.arange(b * batch_size, min((b + 1) * batch_size, n_user), dtype=torch.int32, device=train_dev.device)
        batch = train_dev.batch(users)
        idx = diffusion.rank(model, batch, k, hist=hist_devs[0].csr, hist2=hist_devs[1].csr if len(hist_devs) > 1 else None,
                             steps=sampling_steps, sampling_noise=sampling_noise)
        sums += evaluate_utils.metrics_from_device(idx, users, gt_dev.rowptr, gt_dev.col, topN)
    dist.all_reduce(sums)
    return evaluate_utils.finalize_metrics(sums, n_batches * batch_size if drop_last else n_user)


def evaluate_engine(eng, n_user, batch_size, topN, dist, tail=None):
    """evaluate() through a captured denoise + rank step (StepEngine, train=False, resident mode). tail: callable that
    ranks the users the fixed-size batches leave over (--tst_w_val iterates without drop_last, main.py:174)."""
    n_batches = n_user // batch_size
    dev = eng.dev
    sums = torch.zeros(len(topN), 4, dtype=torch.float64, device=dev)
    for b in range(dist.rank, n_batches, dist.world_size):
        eng.load_users(torch.arange(b * batch_size, (b + 1) * batch_size, dtype=torch.int32, device=dev))
        _, _, s = eng.step()
        sums += s
    n_ranked = n_batches * batch_size
    if tail is not None and n_ranked < n_user:
        if dist.rank == 0:
            sums += tail(n_ranked, n_user)
        n_ranked = n_user
    dist.all_reduce(sums)
    return evaluate_utils.finalize_metrics(sums, n_ranked)


def _gather_ranks(t, dist):
    """[world_size, ...] copy of a per-rank tensor on every rank (checkpoints keep every rank's diffusion state)."""
    if dist.world_size == 1:
        return t.unsqueeze(0).cpu()
    out = [torch.empty_like(t) for _ in range(dist.world_size)]
    torch.distributed.all_gather(out, t.contiguous())
    return torch.stack(out).cpu()


def save_checkpoint(path, epoch, model, optimizer, diffusion, rng, best, dist):
    """Everything a bit-exact continuation needs: weights, AdamW moments and step counters, the
    importance-sampling history and the device-resident RNG epoch OF EVERY RANK (they are rank-local: each rank sees its
    own batches and draws from its own Philox stream), the shuffle generator and the model-selection state.
    Collective: every rank calls it, rank 0 writes."""
    optimizer.flush_lazy()  # row-sparse user-table updates: replay the pending zero-gradient steps before saving
    dev = diffusion.Lt_history.device
    epoch_t = diffusion._epoch if diffusion._epoch is not None else torch.full((1,), -1, dtype=torch.int64, device=dev)
    lt_h, lt_c, ep = (_gather_ranks(t, dist) for t in (diffusion.Lt_history, diffusion.Lt_count, epoch_t))
    if dist.rank != 0:
        return
    torch.save({"epoch": epoch, "model": model.state_dict(), "optimizer": optimizer.state_dict(),
                "optimizer_step_dev": None if optimizer._step_dev is None else int(optimizer._step_dev.item()),
                "world_size": dist.world_size, "Lt_history": lt_h, "Lt_count": lt_c, "rng_epoch": ep,
                "shuffle_rng": rng.bit_generator.state, "best": best}, path)


def load_checkpoint(path, model, optimizer, diffusion, rng, device, dist):
    ck = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(ck["model"])
    model.weights_updated()
    optimizer.load_state_dict(ck["optimizer"])
    if ck["optimizer_step_dev"] is not None:
        optimizer._step_dev = torch.tensor([ck["optimizer_step_dev"]], dtype=torch.int64, device=device)
    r = dist.rank if ck.get("world_size", 1) == dist.world_size else 0  # another world size: every rank starts from rank 0's
    diffusion.Lt_history = ck["Lt_history"][r].to(device)
    diffusion.Lt_count = ck["Lt_count"][r].to(device)
    if int(ck["rng_epoch"][r]) >= 0:
        diffusion._epoch = ck["rng_epoch"][r].reshape(1).to(device)
    rng.bit_generator.state = ck["shuffle_rng"]
    return ck["epoch"], ck["best"]


def main(args):
    dist = dist_utils.init()
    out_path = os.path.join(args.log_name, args.dataset, datetime.now().strftime('%Y%m%d'), args.out_name)
    if dist.rank == 0:
        os.makedirs(out_path, exist_ok=True)
    out_path_file = os.path.join(out_path, 'output_NDCG.txt')
    if not args.debug and dist.rank == 0:
        sys.stdout = open(out_path_file, 'w')
    elif dist.rank != 0:
        sys.stdout = open(os.devnull, 'w')
    print('out_path:', out_path, out_path_file)
    print("args:", args)
    print('random_seed:', args.random_seed)
    torch.manual_seed(args.random_seed)  # the reference prints the seed but never applies it (main.py:123-127)
    np.random.seed(args.random_seed)
    if not torch.cuda.is_available():
        raise RuntimeError("gdmcf_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    device = torch.device("cuda:{}".format(dist.local_rank if dist.world_size > 1 else args.gpu))
    torch.cuda.set_device(device)
    print('device:', device)
    print("Starting time: ", time.strftime('%Y-%m-%d %H:%M:%S', time.localtime(time.time())))

    train_data, valid_y_data, test_y_data, n_user, n_item = load_interactions(args)
    n_rows = n_user if args.n_user <= 0 else min(args.n_user, n_user)
    train_dev = data_utils.DeviceInteractions(train_data[:n_rows], device)
    valid_dev = data_utils.DeviceInteractions(valid_y_data[:n_rows], device)
    test_dev = data_utils.DeviceInteractions(test_y_data[:n_rows], device)
    print('data ready.')

    diffusion, model = build(args, n_rows, n_item, device)
    diffusion.seed = model.seed = args.random_seed + 1000 * dist.rank
    dist.broadcast_parameters(model)
    use_engine = not (args.eager or args.faithful_graph)  # the faithful graph mode builds its edge CSR with host-visible sizes
    optimizer = FusedAdamW(model.parameters(), lr=args.lr, weight_decay=args.weight_decay, modules=[model],
                           capturable=use_engine)
    print("models ready.")
    param_num = sum(p.nelement() for p in model.parameters()) + sum(p.nelement() for p in diffusion.parameters())
    print("Number of all parameters:", param_num)

    topN = eval(args.topN) if isinstance(args.topN, str) else list(args.topN)
    B = args.batch_size
    eval_B = args.eval_batch_size or B
    n_batches = n_rows // B  # drop_last=True (main.py:155)
    best_recall, best_epoch = -100, 0
    best_test_results = None
    rng = np.random.default_rng(args.random_seed)
    start_epoch = 1
    if args.resume:
        last, best = load_checkpoint(args.resume, model, optimizer, diffusion, rng, device, dist)
        best_recall, best_epoch, best_test_results = best
        start_epoch = last + 1
        print('resumed from', args.resume, 'after epoch', last)
    # test-time input rows: the training rows, or training + validation rows with --tst_w_val (main.py:172-175,354-358)
    tv_dev = None
    if args.tst_w_val:
        tv_dev = data_utils.DeviceInteractions((train_data[:n_rows] + valid_y_data[:n_rows]).astype(bool).astype('float64'), device)
    train_eng = valid_eng = test_eng = None
    if use_engine:
        # the loop body (main.py:331-351) and evaluate() (main.py:267-310) as captured steps; the warm-up steps the capture
        # needs are rolled back (preserve_state), so the run starts from the same state as the eager loop
        if n_batches > 0:
            train_eng = StepEngine(model, diffusion, optimizer, dist, batch_size=B, n_item=n_item, topk=topN[-1], topN=topN,
                                   cap_train_nnz=1, cap_gt_nnz=1, reweight=args.reweight, rank=False, nccl_sms=args.nccl_sms)
            train_eng.bind_resident(train_dev)
            train_eng.load_users(torch.arange(B, dtype=torch.int32, device=device))
            train_eng.capture(warmup=2, preserve_state=True)
        if n_rows // eval_B > 0:
            def mk():
                return StepEngine(model, diffusion, None, dist_utils.Dist(), batch_size=eval_B, n_item=n_item, topk=topN[-1],
                                  topN=topN, cap_train_nnz=1, cap_gt_nnz=1, train=False, sampling_steps=args.sampling_steps,
                                  sampling_noise=args.sampling_noise)
            valid_eng, test_eng = mk(), mk()
            valid_eng.bind_resident(train_dev, gt_dev=valid_dev, hist_dev=train_dev)
            if args.tst_w_val:
                test_eng.bind_resident(tv_dev, gt_dev=test_dev, hist_dev=tv_dev)
            else:
                test_eng.bind_resident(train_dev, gt_dev=test_dev, hist_dev=train_dev, hist2_dev=valid_dev)
            for e in (valid_eng, test_eng):
                e.load_users(torch.arange(eval_B, dtype=torch.int32, device=device))
                e.capture(warmup=1, preserve_state=True)

    def tail_rows(x_dev, gt_dev, hists):
        def run(lo, hi):  # users the fixed-size batches leave over: ranked call by call
            users = torch.arange(lo, hi, dtype=torch.int32, device=device)
            idx = diffusion.rank(model, x_dev.batch(users), topN[-1], hist=hists[0].csr,
                                 hist2=hists[1].csr if len(hists) > 1 else None, steps=args.sampling_steps,
                                 sampling_noise=args.sampling_noise)
            return evaluate_utils.metrics_from_device(idx, users, gt_dev.rowptr, gt_dev.col, topN)
        return run

    def run_evaluation():
        model.eval()
        if train_eng is not None:
            train_eng.flush()  # row-sparse user-table updates: every user's row must be current before all users are ranked
        if valid_eng is not None:
            valid = evaluate_engine(valid_eng, n_rows, eval_B, topN, dist)
        else:
            valid = evaluate(diffusion, model, train_dev, valid_dev, [train_dev], n_rows, eval_B, topN, args.sampling_steps, dist,
                             sampling_noise=args.sampling_noise)
        if args.tst_w_val:
            if test_eng is not None:
                test = evaluate_engine(test_eng, n_rows, eval_B, topN, dist, tail=tail_rows(tv_dev, test_dev, [tv_dev]))
            else:
                test = evaluate(diffusion, model, tv_dev, test_dev, [tv_dev], n_rows, eval_B, topN, args.sampling_steps, dist,
                                drop_last=False, sampling_noise=args.sampling_noise)
        elif test_eng is not None:
            test = evaluate_engine(test_eng, n_rows, eval_B, topN, dist)
        else:
            test = evaluate(diffusion, model, train_dev, test_dev, [train_dev, valid_dev], n_rows, eval_B, topN,
                            args.sampling_steps, dist, sampling_noise=args.sampling_noise)
        return valid, test

    print("Start training...")
    stats = {"train_s": [], "eval_s": [], "users_per_epoch": (n_batches - n_batches % dist.world_size) * B,
             "users_per_eval": 2 * (n_rows // eval_B) * eval_B, "engine": use_engine}
    main.last_stats = stats  # wall-clock seconds per epoch of training / per evaluation (valid + test), for callers and tests
    for epoch in range(start_epoch, args.epochs + 1):
        if epoch - best_epoch >= 200:
            print('-' * 18)
            print('Exiting from training early')
            break
        model.train()
        start_time = time.time()
        total_loss = torch.zeros((), dtype=torch.float64, device=device)
        perm = torch.from_numpy(rng.permutation(n_rows).astype(np.int32)).to(device)  # shuffle=True
        # data parallel: G consecutive logical batches form one optimizer step
        for b0 in range(0, n_batches - n_batches % dist.world_size, dist.world_size):
            b = b0 + dist.rank
            if train_eng is not None:
                train_eng.load_users(perm[b * B:(b + 1) * B])
                loss, _, _ = train_eng.step()
                total_loss += loss
                continue
            batch = train_dev.batch(perm[b * B:(b + 1) * B])
            optimizer.zero_grad()
            losses = diffusion.training_losses(model, batch, args.reweight, index=batch.users)
            loss = losses["loss"].mean()
            total_loss += loss.detach()
            loss.backward()
            dist.all_reduce_gradients(model)
            optimizer.step(grad_scale=1.0 / dist.world_size)

        torch.cuda.synchronize(device)
        train_s = time.time() - start_time  # the epoch's training steps alone (evaluation below is timed separately)
        if epoch % args.eval_every == 0:
            t_eval = time.time()
            valid_results, test_results = run_evaluation()
            torch.cuda.synchronize(device)
            stats["eval_s"].append(time.time() - t_eval)
            evaluate_utils.print_results(None, valid_results, test_results)
            sys.stdout.flush()
            if valid_results[2][1] > best_recall:  # NDCG@topN[1] despite the name (main.py:362-363)
                best_recall, best_epoch = test_results[2][1], epoch
                best_test_results = test_results
                if dist.rank == 0:
                    model_path = os.path.join(out_path, 'model.pth')
                    print('model_path:', model_path)
                    torch.save(model, model_path)
        if args.checkpoint_every and epoch % args.checkpoint_every == 0:
            if train_eng is not None:
                train_eng.flush()  # lazily updated user rows; with several ranks also the fp32 masters gathered as bf16 operands
                train_eng.gather_optimizer_state()  # row-sharded AdamW: the moments of the other ranks' row blocks
            save_checkpoint(os.path.join(out_path, 'checkpoint.pt'), epoch, model, optimizer, diffusion, rng,
                            (best_recall, best_epoch, best_test_results), dist)
        stats["train_s"].append(train_s)
        print("Runing Epoch {:03d} ".format(epoch) + 'train loss {:.4f}'.format(float(total_loss)) + " costs " +
              time.strftime("%H: %M: %S", time.gmtime(time.time() - start_time)))
        print('---' * 18)

    print('===' * 18)
    print("End. Best Epoch {:03d} ".format(best_epoch))
    evaluate_utils.print_results(None, None, best_test_results)
    if train_eng is not None:
        train_eng.flush()
    main.last_model = model  # handle for callers / tests (the reference keeps everything local to main())
    print("End time: ", time.strftime('%Y-%m-%d %H:%M:%S', time.localtime(time.time())))
    dist.shutdown()
    return best_test_results


if __name__ == '__main__':
    main(parse_args())
