"""TEST INFRASTRUCTURE — generates tests/golden/*.npz + kat.json by executing the UNMODIFIED reference modules
(/root/reference, imported through oracle/ref_harness.py) on small seeded inputs. Run in the authoring
container only:  python oracle/make_golden.py
The fixtures pin oracle/gdmcf_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
B, I, U, D, E, T = 12, 150, 40, 32, 10, 5


def np_(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def sd_np(model, prefix):
    return {f"{prefix}{k}": np_(v) for k, v in model.state_dict().items()}


def synth_interactions(seed=0):
    rng = np.random.default_rng(seed)
    dense = (rng.random((U, I)) < 0.08).astype(np.float32)
    dense[:, 0] = 1.0  # a hub item
    dense[3] = 0.0
    dense[3, 5] = 1.0  # a nearly-empty user
    test = ((rng.random((U, I)) < 0.03) & (dense == 0)).astype(np.float32)
    test[7] = 0.0  # a user with empty ground truth (counts in the denominator, evaluate_utils.py:17,47)
    return dense, test


def split_draws(rec, x0):
    """Turn the reference's recorded multinomial outcome of apply_noise into an injectable u_keep."""
    n = x0.numel()
    m = [r for r in rec.multinomial if r.numel() == n]
    assert len(m) >= 1
    kept = (m[0].reshape(x0.shape) == x0.long())
    return torch.where(kept, 0.0, 1.0)


def main():
    os.makedirs(GOLD, exist_ok=True)
    gd, dnn, ev, du = H.import_reference()
    kat = {}

    # ---------------------------------------------------------------- schedules / KATs
    sched = {}
    for steps, ns in [(5, 0.01), (5, 0.0001), (10, 0.01), (50, 0.01), (100, 0.01), (5, 0.1)]:
        diff = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", ns, 0.001, 0.01, steps, "cpu",
                                            discrete=0.9995, CatOneHot=True, args=H.make_args())
        ts = torch.arange(steps)
        w = torch.where(ts == 0, 1.0, diff.SNR(ts - 1) - diff.SNR(ts))
        sched[f"steps{steps}_ns{ns}"] = {k: np_(getattr(diff, k)).tolist() for k in (
            "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
            "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")}
        sched[f"steps{steps}_ns{ns}"]["reweight"] = np_(w).tolist()
    kat["schedule"] = sched
    kat["Qt_bar"] = {str(a): np_(diff.get_Qt_bar(torch.tensor([a], dtype=torch.float32)))[0].tolist() for a in (0.0, 4 / 400)}
    kat["timestep_embedding"] = {str(t): np_(dnn.timestep_embedding(torch.tensor([t]), 10))[0].tolist() for t in (0, 1, 4, 99)}
    GT = [[1, 5, 7], [], [2], [9, 3]]
    pred = [[5, 0, 7, 2, 4], [1, 2, 3, 4, 5], [0, 1, 3, 4, 5], [3, 9, 1, 2, 0]]
    kat["topn"] = {"GT": GT, "pred": pred, "topN": [1, 3, 5], "out": [list(x) for x in ev.computeTopNAccuracy(GT, pred, [1, 3, 5])]}
    with open(os.path.join(GOLD, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    dense, test = synth_interactions()
    x_all = torch.from_numpy(dense)
    index = torch.tensor([0, 3, 5, 7, 8, 11, 13, 20, 21, 30, 38, 39])
    x0 = x_all[index]
    args = H.make_args()

    def new_diffusion(steps=T, ns=0.01, cat=True, index_in=True):
        d = gd.GaussianDiffusionDiscrete(gd.ModelMeanType.START_X, "linear-var", ns, 0.001, 0.01, steps, "cpu",
                                         discrete=0.9995, CatOneHot=cat, args=args)
        d.indexIn = index_in
        return d

    # ---------------------------------------------------------------- DNN backbone
    torch.manual_seed(0)
    model = dnn.DNN([I, D], [D, I], E, time_type="cat", norm=False)
    out = {"x0": np_(x0), "index": np_(index)}
    out.update(sd_np(model, "sd."))
    model.eval()
    ts = torch.tensor([0, 1, 2, 3, 4, 4, 3, 2, 1, 0, 2, 4])
    xt = x0 + 0.01 * torch.randn(B, I)
    out["fwd_ts"], out["fwd_x"] = np_(ts), np_(xt)
    with torch.no_grad():
        out["fwd_eval"] = np_(model(xt, ts))
    diff = new_diffusion(cat=False, index_in=False)
    with torch.no_grad():
        out["p_sample_s0"] = np_(diff.p_sample(model, x0.clone(), 0, False, index=index))
    # training: 14 steps so that importance sampling switches on (Lt_count reaches 10 for every t)
    model.train()
    diff = new_diffusion(cat=False, index_in=False)
    steps_rec = []
    torch.manual_seed(1)
    for it in range(14):
        model.zero_grad()
        with H.record_rng(gd) as rec:
            terms = diff.training_losses(model, x0, True, index=index)
        loss = terms["loss"]
        loss.mean().backward()
        if len(rec.randint):
            ts_used = rec.randint[-1]
        else:
            ts_used = [r for r in rec.multinomial if r.numel() == B][-1]
        steps_rec.append({"ts": np_(ts_used), "noise": np_(rec.randn[0]), "keep_x": np_(rec.dropout[0]),
                          "loss": np_(loss), "ready": bool(len(rec.randint) == 0)})
        if it in (0, 13):
            for k, p in model.named_parameters():
                out[f"train.grad{it}.{k}"] = np_(p.grad) if p.grad is not None else np.zeros(0, dtype=np.float32)
    for k in steps_rec[0]:
        out[f"train.{k}"] = np.stack([s[k] for s in steps_rec])
    out["train.Lt_history"], out["train.Lt_count"] = np_(diff.Lt_history), np_(diff.Lt_count)
    np.savez_compressed(os.path.join(GOLD, "dnn_backbone.npz"), **out)

    # ---------------------------------------------------------------- GDMCF backbone
    torch.manual_seed(2)
    model = dnn.DNNOneHotEmbeddingGCN([I, D], [D, I], E, time_type="cat", norm=False, item_num=I, user_num=U, args=args)
    with torch.no_grad():
        model.sumW.fill_(0.8)  # exercise the GCN mix (1.0 at init would hide it)
        model.gcn_model.conv1.bias.normal_(0, 0.01)
        model.gcn_model.conv2.bias.normal_(0, 0.01)
    out = {"x0": np_(x0), "index": np_(index), "test": test[np_(index)], "train_all": dense, "test_all": test}
    out.update(sd_np(model, "sd."))
    model.eval()
    x_U = F.one_hot(x0.long(), 2).float()
    out["fwd_ts"], out["fwd_x"] = np_(ts), np_(xt)
    with torch.no_grad():
        o_edges = model(xt, ts, x_U, index=index, graph=x_U.long())
        o_noedge = model(xt, ts, x_U, index=index, graph=torch.zeros_like(x_U).long())
    assert torch.equal(o_edges, o_noedge), "user rows must not depend on the edge set (SURVEY.md §0)"
    out["fwd_eval"] = np_(o_edges)
    diff = new_diffusion()
    with torch.no_grad():
        p0 = diff.p_sample(model, x0.clone(), 0, False, index=index)
    out["p_sample_s0"] = np_(p0)
    torch.manual_seed(3)
    with torch.no_grad(), H.record_rng(gd) as rec:
        p2 = diff.p_sample(model, x0.clone(), 2, False, index=index)
    out["p_sample_s2"], out["p_sample_s2.noise"] = np_(p2), np_(rec.randn[0])
    out["p_sample_s2.u_keep"] = np_(split_draws(rec, x0))
    # steps sweep on the same weights (config 4)
    # (the reference's graph-noise block needs a = t/B <= 1, gaussian_diffusion.py:775: steps <= batch size)
    for steps in (10, 12):
        d2 = new_diffusion(steps=steps)
        with torch.no_grad():
            out[f"p_sample_s0_T{steps}"] = np_(d2.p_sample(model, x0.clone(), 0, False, index=index))
    # evaluate closure (main.py:267-310) on this batch: mask train history, top-K, metrics
    topN = [5, 10, 20]
    pred = p0.clone()
    his = torch.from_numpy(dense[np_(index)])
    pred[his.nonzero(as_tuple=True)] = -np.inf
    vals, idx = torch.topk(pred, topN[-1])
    target = [np.nonzero(test[u])[0].tolist() for u in np_(index)]
    out["eval.topk_idx"], out["eval.topk_val"] = np_(idx), np_(vals)
    out["eval.metrics"] = np.array(ev.computeTopNAccuracy(target, idx.tolist(), topN))
    out["eval.topN"] = np.array(topN)
    # training with recorded draws
    model.train()
    diff = new_diffusion()
    steps_rec = []
    torch.manual_seed(4)
    for it in range(14):
        model.zero_grad()
        with H.record_rng(gd) as rec:
            terms = diff.training_losses(model, x0, True, index=index)
        loss = terms["loss"]
        loss.mean().backward()
        if len(rec.randint):
            ts1, ts2 = rec.randint[0], rec.randint[1]
        else:
            mm = [r for r in rec.multinomial if r.numel() == B]
            ts1, ts2 = mm[0], mm[1]
        steps_rec.append({"ts_discrete": np_(ts1), "ts": np_(ts2), "noise": np_(rec.randn[0]),
                          "u_keep": np_(split_draws(rec, x0)), "keep_x": np_(rec.dropout[0]),
                          "keep_xU": np_(rec.dropout[1]), "loss": np_(loss)})
        if it in (0, 13):
            for k, p in model.named_parameters():
                out[f"train.grad{it}.{k}"] = np_(p.grad) if p.grad is not None else np.zeros(0, dtype=np.float32)
    for k in steps_rec[0]:
        out[f"train.{k}"] = np.stack([s[k] for s in steps_rec])
    out["train.Lt_history"], out["train.Lt_count"] = np_(diff.Lt_history), np_(diff.Lt_count)
    # nt_xent on its own
    z1, z2 = torch.randn(B, D) * 0.3, torch.randn(B, D) * 0.3
    out["ntxent.z1"], out["ntxent.z2"], out["ntxent.out"] = np_(z1), np_(z2), np_(dnn.nt_xent_loss(z1, z2))
    np.savez_compressed(os.path.join(GOLD, "gdmcf_backbone.npz"), **out)

    # ---------------------------------------------------------------- apply_noise keep statistics
    torch.manual_seed(5)
    diff = new_diffusion()
    big = (torch.rand(400, 2000) < 0.05).float()
    t4 = torch.full((400,), 4)
    xn = diff.apply_noise(t4, F.one_hot(big.long(), 2).float()) & F.one_hot(big.long(), 2)
    keep1 = xn[..., 1].sum().item() / big.sum().item()
    keep0 = xn[..., 0].sum().item() / (1 - big).sum().item()
    kat["apply_noise_keep_rate_B400_t4"] = {"ones": keep1, "zeros": keep0}
    with open(os.path.join(GOLD, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---------------------------------------------------------------- LightGCN propagation (lightGCN.py:129-203)
    nu, ni, dl, K = 60, 45, 64, 3
    rng = np.random.default_rng(6)
    pairs = np.unique(np.stack([rng.integers(0, nu, 500), rng.integers(0, ni, 500)], 1), axis=0)
    pairs = np.concatenate([pairs, np.stack([np.arange(nu), np.zeros(nu, dtype=np.int64)], 1)])  # hub item 0
    pairs = np.unique(pairs, axis=0)
    df = {"user_id_idx": pairs[:, 0], "item_id_idx": pairs[:, 1]}  # the class only indexes these two columns
    LightGCN = H.load_lightgcn_class(nu, ni)
    torch.manual_seed(7)
    lg = LightGCN(df, nu, ni, K, dl)
    A = lg.norm_adj_mat_sparse_tensor.coalesce()
    with torch.no_grad():
        fu, fi, iu, ii = lg.propagate_through_layers()
    np.savez_compressed(os.path.join(GOLD, "lightgcn.npz"), pairs=pairs, n_users=nu, n_items=ni, n_layers=K,
                        E0=np_(lg.E0.weight), A_indices=np_(A.indices()), A_values=np_(A.values()),
                        final_user=np_(fu), final_item=np_(fi))
    print("golden fixtures written to", GOLD, {f: os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD)})


if __name__ == "__main__":
    main()
