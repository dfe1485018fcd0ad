"""Host-side work plan of the bf16 LightGCN propagation (kernels.lightgcn_plan_bf16; lightGCN.py:180-194 is the op it
serves): no GPU needed — the plan is built by the library's host function + numpy. Checks the invariants the kernel relies
on: every neighbour of every row is covered exactly once (hub pieces + whole rows), lists are hot-first with the hot
neighbours encoded as shared-memory slots, the per-warp row lists partition the whole-row items, and the estimated work is
balanced across warp slots."""
import numpy as np
import torch


def _graph(n=3000, seed=0):
    rng = np.random.default_rng(seed)
    deg = np.minimum(rng.zipf(1.6, n) + 1, n // 2)
    deg[:3] = [900, 700, 400]                      # hub rows: split into pieces
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = np.concatenate([rng.choice(n, d, replace=False) for d in deg]).astype(np.int64)
    return rowptr, col, deg


def test_plan_covers_every_neighbour_once_and_is_balanced(lib):
    from gdmcf_b200 import kernels as K
    rowptr, col, deg = _graph()
    n = len(deg)
    plan = K.lightgcn_plan_bf16(torch.from_numpy(rowptr.astype(np.int32)), torch.from_numpy(col.astype(np.int32)), device="cpu")
    items, mids = plan.items.numpy()[: plan.n_items], plan.mids.numpy()[: plan.n_items]
    enc = plan.col.numpy().view(np.uint32)
    hot_rows, longs = plan.hot_rows.numpy()[: plan.n_hot], plan.long_rows.numpy()[: plan.n_long]
    # -- pieces first, then whole rows; every nnz position belongs to exactly one item
    assert (items[: plan.n_pieces, 3] >= 0).all() and (items[plan.n_pieces:, 3] < 0).all()
    cover = np.zeros(len(col), dtype=np.int32)
    for (_, b, e, _) in items:
        cover[b:e] += 1
    assert (cover == 1).all()
    # -- whole-row items span exactly their row; pieces of hub row li stay inside that row and carry -(li + 1)
    for row, b, e, slot in items[plan.n_pieces:]:
        assert b == rowptr[row] and e == rowptr[row + 1]
    for x, b, e, slot in items[: plan.n_pieces]:
        row = longs[-x - 1, 0]
        assert rowptr[row] <= b < e <= rowptr[row + 1]
    assert plan.n_long >= 3 and {0, 1, 2} <= {int(r) for r in longs[:, 0]}  # the three hub rows are cut into pieces
    # -- hot-first lists: [begin, mid) are slots of the hot set, [mid, end) global ids; together = the original neighbours
    slot_row = {s: int(r) for s, r in enumerate(hot_rows)}
    for (x, b, e, slot), mid in zip(items, mids):
        assert b <= mid <= e
        assert (enc[b:mid] & 0x80000000).all() and not (enc[mid:e] & 0x80000000).any()
    for row in range(0, n, 97):
        b, e = rowptr[row], rowptr[row + 1]
        got = sorted(slot_row[int(v & 0x7FFFFFFF)] if v & 0x80000000 else int(v) for v in enc[b:e])
        assert got == sorted(col[b:e].tolist())
    cnt = np.bincount(col, minlength=n)
    assert cnt[hot_rows].min() >= np.sort(cnt)[::-1][plan.n_hot - 1]  # the hot set = the most gathered rows
    # -- per-warp row lists: a partition of the whole-row items, pairs of similar cost, balanced totals
    wp = plan.warp_ptr.numpy()
    assert wp[0] == 0 and wp[-1] == plan.n_items - plan.n_pieces and (np.diff(wp) >= 0).all() and len(wp) == plan.n_slots + 1
    cost = np.ceil((items[:, 2] - mids) / 8) + 0.35 * np.ceil((mids - items[:, 1]) / 8)
    rows_cost = cost[plan.n_pieces:]
    load = np.zeros(plan.n_slots)
    for w in range(plan.n_slots):
        seg = rows_cost[wp[w]:wp[w + 1]]
        load[w] = sum(max(seg[i:i + 2]) + 1.0 for i in range(0, len(seg), 2))  # two rows per warp: the heavier sets the trips
    pc = np.ceil((items[: plan.n_pieces, 2] - mids[: plan.n_pieces]) / 16) + 0.35 * np.ceil((mids[: plan.n_pieces] - items[: plan.n_pieces, 1]) / 16) + 1.5
    np.add.at(load, np.arange(plan.n_pieces) % plan.n_slots, pc)
    busy = load[load > 0]
    assert busy.max() <= busy.mean() + max(rows_cost.max(), pc.max() if plan.n_pieces else 0) + 1.0  # LPT: within one item of the mean
