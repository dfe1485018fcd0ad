"""GPU probe: correctness diagnostics + CUDA-event timings of the individual kernels at the BASELINE shapes.
Writes gpurun_out/probe_<tag>.json. Usage: python tools/gpu_probe.py [tag] [sections...]"""
import json
import math
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdmcf_b200 import kernels as K  # noqa: E402

OUT = {}


def timeit(fn, warmup=3, iters=10, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return {"ms_min": ts[0], "ms_med": ts[len(ts) // 2]}


def bf16_op(rows, cols, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.zeros(rows, K.round_up(cols, 64), dtype=torch.bfloat16, device="cuda")
    t[:, :cols] = (torch.randn(rows, cols, generator=g, device="cuda") * scale).to(torch.bfloat16)
    return t


def section(name):
    def deco(fn):
        def run():
            t0 = time.time()
            try:
                OUT[name] = fn()
            except Exception as ex:  # noqa: BLE001
                OUT[name] = {"error": repr(ex), "trace": traceback.format_exc()[-1500:]}
            OUT[name + "_wall_s"] = round(time.time() - t0, 2)
            print(name, json.dumps(OUT[name])[:600], flush=True)
        run.__name__ = name
        return run
    return deco


@section("gemm_diag")
def gemm_diag():
    res = {}
    for (m, n, k, splits) in [(128, 256, 64, 1), (128, 256, 256, 1), (128, 128, 128, 1), (256, 512, 512, 1), (400, 1000, 3000, 1),
                              (128, 256, 1024, 4)]:
        a, b = bf16_op(m, k, 1), bf16_op(n, k, 2)
        out = torch.full((m, K.round_up(n, 4)), float("nan"), device="cuda")
        K.gemm([a], [b], m, n, [k], out_f32=out, splits=splits)
        torch.cuda.synchronize()
        ref = a[:, :k].float() @ b[:, :k].float().t()
        got = out[:, :n]
        err = (got - ref).abs()
        bad = torch.isnan(got).sum().item()
        info = {"max_err": float(err.nan_to_num(1e9).max()), "ref_max": float(ref.abs().max()), "nan": bad}
        if info["max_err"] > 1e-3 * info["ref_max"]:
            # where is it wrong? per 32x32 block error map (first 8x8 blocks)
            e2 = err.nan_to_num(1e9)[: min(m, 256), : min(n, 256)]
            bm = e2.reshape(e2.shape[0] // 32, 32, e2.shape[1] // 32, 32).amax(dim=(1, 3))
            info["blockmap"] = [[round(float(x), 3) for x in row] for row in bm.cpu()]
            info["got00"] = [float(x) for x in got[0, :8].cpu()]
            info["ref00"] = [float(x) for x in ref[0, :8].cpu()]
            # is it a K-subset? compare against partial-K references
            for kk in (16, 32, 64, 128):
                if kk <= k:
                    rp = a[:, :kk].float() @ b[:, :kk].float().t()
                    info[f"err_vs_k{kk}"] = float((got - rp).abs().nan_to_num(1e9).max())
        res[f"{m}x{n}x{k}s{splits}"] = info
    return res


@section("gemm_perf")
def gemm_perf():
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = {
        "enc_yelp_B512": (512, 1000, 34395, None),
        "enc_yelp_B400": (400, 1000, 34395, None),
        "score_yelp_B512": (512, 34395, 3000, 1),
        "score_yelp_B400": (400, 34395, 3000, 1),
        "score_yelp_B2048": (2048, 34395, 3000, 1),
        "enc_amz_B512": (512, 1000, 94949, None),
        "score_amz_B512": (512, 94949, 3000, 1),
        "square_8192": (8192, 8192, 8192, 1),
        "gcn1_B512": (512, 512, 3000, None),
        "gcn2_B512": (512, 3000, 512, None),
    }
    for name, (m, n, k, splits) in shapes.items():
        a, b = bf16_op(m, k, 1, 0.05), bf16_op(n, k, 2, 0.05)
        out = torch.empty(m, K.round_up(n, 4), device="cuda")
        fn = lambda: K.gemm([a], [b], m, n, [k], out_f32=out, splits=splits)  # noqa: E731
        t = timeit(fn, flush=flush)
        flops = 2.0 * m * n * k
        t["tflops_med"] = flops / (t["ms_med"] * 1e-3) / 1e12
        t["tflops_best"] = flops / (t["ms_min"] * 1e-3) / 1e12
        t["splits"] = K.load().gdmcf_gemm_auto_splits(m, n, K.round_up(k, 64)) if splits is None else splits
        # torch (cuBLAS) for reference on the same operands
        af, bfm = a[:, :k].contiguous(), b[:, :k].contiguous()
        t["cublas_ms_med"] = timeit(lambda: torch.matmul(af, bfm.t()), flush=flush)["ms_med"]
        res[name] = t
        del a, b, out, af, bfm
    return res


def yelp_graph(U=54574, I=34395, nnz=1402736, seed=0, train_frac=0.7):
    rng = np.random.default_rng(seed)
    mean_deg = nnz / U
    deg = rng.lognormal(math.log(mean_deg) - 0.5, 1.0, U)
    deg = np.clip(deg * (nnz / deg.sum()), 5, I // 4).astype(np.int64)
    p = 1.0 / np.arange(1, I + 1)
    p /= p.sum()
    perm = rng.permutation(I)
    users = np.repeat(np.arange(U), deg)
    items = perm[rng.choice(I, size=users.shape[0], p=p)]
    key = np.unique(users.astype(np.int64) * I + items)
    users, items = key // I, key % I
    keep = rng.random(users.shape[0]) < train_frac
    return users[keep], items[keep]


def csr(rows, cols, n_rows):
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32)


@section("spmm_perf")
def spmm_perf():
    res = {}
    U, I = 54574, 34395
    u, i = yelp_graph()
    r_rowptr, r_col = csr(u, i, U)
    rt_rowptr, rt_col = csr(i, u, I)
    dev = "cuda"
    rowptr, col, val = K.build_norm_adj(*(torch.from_numpy(x).to(dev) for x in (r_rowptr, r_col, rt_rowptr, rt_col)), U, I)
    N, nnz2 = U + I, col.numel()
    res["N"], res["nnz"] = N, nnz2
    rp = rowptr.cpu().numpy()
    res["max_deg"] = int(np.diff(rp).max())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    E0 = torch.randn(N, 64, device=dev) * 0.1
    # correctness vs torch.sparse
    crow = rowptr.long()
    A = torch.sparse_csr_tensor(crow, col.long(), val, size=(N, N))
    ref = E0
    layers = [E0]
    for _ in range(3):
        ref = torch.sparse.mm(A, ref)
        layers.append(ref)
    ref_mean = torch.stack(layers).mean(0)
    for chunk in (64, 128, 256, 512):
        plan = K.spmm_plan(rp, chunk=chunk)
        out = K.lightgcn_propagate(plan, col, val, E0, 3)
        torch.cuda.synchronize()
        err = float((out - ref_mean).abs().max())
        work = (torch.empty_like(E0), torch.empty_like(E0), torch.empty(max(plan.n_slots, 1), 64, device=dev))
        o2 = torch.empty_like(E0)
        t = timeit(lambda: K.lightgcn_propagate(plan, col, val, E0, 3, out=o2, work=work), flush=flush, iters=20)
        d, s = 64, 4
        bytes_layer = nnz2 * (4 + s) + (N + 1) * 4 + 2 * N * d * s
        t["GBs_med"] = 3 * bytes_layer / (t["ms_med"] * 1e-3) / 1e9
        t["GBs_best"] = 3 * bytes_layer / (t["ms_min"] * 1e-3) / 1e9
        t["err"] = err
        t["n_items"], t["n_long"], t["n_slots"] = plan.n_items, plan.n_long, plan.n_slots
        res[f"chunk{chunk}"] = t
    t = timeit(lambda: torch.sparse.mm(A, E0), flush=flush)
    res["torch_sparse_mm_layer_ms"] = t["ms_med"]
    return res


@section("topk_perf")
def topk_perf():
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, (B, I, k) in {"yelp_B400_k100": (400, 34395, 100), "yelp_B2048_k20": (2048, 34395, 20), "amz_B400_k100": (400, 94949, 100)}.items():
        ld = K.round_up(I, 4)
        scores = torch.randn(B, ld, device="cuda")
        deg = 20
        rowptr = torch.arange(0, (B + 1) * deg, deg, dtype=torch.int32, device="cuda")
        hcol = torch.randint(0, I, (B * deg,), dtype=torch.int32, device="cuda")
        t = timeit(lambda: K.mask_topk(scores, B, I, k, hist=(rowptr, hcol)), flush=flush)
        t["GBs_med"] = B * I * 4 / (t["ms_med"] * 1e-3) / 1e9
        t["torch_topk_ms"] = timeit(lambda: torch.topk(scores[:, :I], k), flush=flush)["ms_med"]
        res[name] = t
    return res


@section("elementwise_perf")
def elementwise_perf():
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, I = 400, 34395
    x0 = (torch.rand(B, K.round_up(I, 4), device="cuda") < 0.001).float()
    a = K.Bf16Mat.empty(B, I, "cuda")
    xt = torch.empty(B, K.round_up(I, 4), device="cuda")
    ts = torch.randint(0, 5, (B,), dtype=torch.int32, device="cuda")
    sa = torch.rand(5, device="cuda")
    t = timeit(lambda: K.qsample_dropout(x0, B, I, a, row_t=ts, sqrt_ab=sa, sqrt_1mab=sa, dropout_p=0.5, seed=1, xt_out=xt), flush=flush)
    t["GBs_med"] = B * I * (4 + 4 + 2) / (t["ms_med"] * 1e-3) / 1e9
    res["qsample_dropout"] = t
    out = torch.zeros(B, K.round_up(2 * I, 64), dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: K.onehot_noise(x0, B, I, out, ts=ts, dropout_p=0.5, seed=1), flush=flush)
    t["GBs_med"] = B * I * (4 + 4) / (t["ms_med"] * 1e-3) / 1e9
    res["onehot_noise"] = t
    n = 68_000_000
    p, g, m, v = (torch.randn(n, device="cuda") for _ in range(4))
    v.abs_()
    t = timeit(lambda: K.adamw_fused(p, g, m, v, lr=1e-5, step=3), flush=flush)
    t["GBs_med"] = n * 28 / (t["ms_med"] * 1e-3) / 1e9
    res["adamw_68M"] = t
    W = torch.randn(1000, 34405, device="cuda")
    wb = K.Bf16Mat.empty(1000, 34405, "cuda")
    t = timeit(lambda: K.cast_bf16(W, out=wb), flush=flush)
    t["GBs_med"] = 1000 * 34405 * 6 / (t["ms_med"] * 1e-3) / 1e9
    res["cast_bf16_W1"] = t
    return res


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    wanted = sys.argv[2:] or ["elementwise_perf", "topk_perf", "spmm_perf", "gemm_diag", "gemm_perf"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    OUT["device"] = torch.cuda.get_device_name(0)
    for name in wanted:
        globals()[name]()
        with open(os.path.join(ROOT, "gpurun_out", f"probe_{tag}.json"), "w") as f:
            json.dump(OUT, f, indent=1)
