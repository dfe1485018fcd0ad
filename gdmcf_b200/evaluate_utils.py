"""Ranking metrics — mirror of the reference's evaluate_utils.py (computeTopNAccuracy :6-52, print_results :54-69).

`computeTopNAccuracy(GroundTruth, predictedIndices, topN)` keeps the reference's list-based signature and 4-dp
rounding; the counting runs on the device (gdmcf_topn_metrics: one warp per user, binary search in the sorted
ground-truth row, sums accumulated in the reference's order) instead of the O(users * topN * len(GT)) Python loop.
`metrics_from_device` is the zero-copy form used by main.py (top-K indices and ground-truth CSR already on the GPU).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import kernels as K


_topn_cache: dict = {}


def _topn_dev(topN: Sequence[int], device) -> torch.Tensor:
    """Cut-offs as a cached device tensor (a fresh torch.tensor(...) is a pageable H2D copy: illegal in graph capture)."""
    key = (tuple(int(x) for x in topN), str(device))
    t = _topn_cache.get(key)
    if t is None:
        t = torch.tensor(list(key[0]), dtype=torch.int32, device=device)
        _topn_cache[key] = t
    return t


def metrics_from_device(topk_idx: torch.Tensor, users, gt_rowptr: torch.Tensor, gt_col: torch.Tensor, topN: Sequence[int]):
    """Per-cutoff sums [len(topN), 4] (precision, recall, NDCG, MRR numerators) as a float64 device tensor."""
    assert max(topN) <= topk_idx.shape[1], "topN[-1] exceeds the number of ranked items"
    topn_dev = _topn_dev(topN, topk_idx.device)
    stats = K.topn_metrics(topk_idx, users, gt_rowptr, gt_col, topn_dev, len(topN))
    return K.colsum_f64(stats).reshape(len(topN), 4)


def finalize_metrics(sums: torch.Tensor, n_users: int):
    """Divide by the number of ranked users (including those with empty ground truth, evaluate_utils.py:47-50)."""
    s = sums.cpu().numpy()
    cols = [[round(float(s[j, q]) / n_users, 4) for j in range(s.shape[0])] for q in range(4)]
    return cols[0], cols[1], cols[2], cols[3]


def computeTopNAccuracy(GroundTruth: List[List[int]], predictedIndices, topN: List[int]):
    """Reference signature: python lists in, (precision, recall, NDCG, MRR) lists out."""
    if not torch.cuda.is_available():
        raise RuntimeError("gdmcf_b200.evaluate_utils needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda")
    n = len(predictedIndices)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum([len(g) for g in GroundTruth[:n]])
    col = np.concatenate([np.sort(np.asarray(g, dtype=np.int32)) for g in GroundTruth[:n]] + [np.zeros(0, np.int32)])
    pred = torch.as_tensor(np.asarray(predictedIndices, dtype=np.int32), device=dev)
    gt_col = torch.from_numpy(col if col.size else np.zeros(1, np.int32)).to(dev)
    sums = metrics_from_device(pred, None, torch.from_numpy(rowptr).to(dev), gt_col, topN)
    return finalize_metrics(sums, n)


def print_results(loss, valid_result, test_result):
    """output the evaluation results (format identical to evaluate_utils.py:54-69)."""
    if loss is not None:
        print("[Train]: loss: {:.4f}".format(loss))
    for tag, res in (("Valid", valid_result), ("Test", test_result)):
        if res is not None:
            print("[{}]: Precision: {} Recall: {} NDCG: {} MRR: {}".format(
                tag, '-'.join([str(x) for x in res[0]]), '-'.join([str(x) for x in res[1]]),
                '-'.join([str(x) for x in res[2]]), '-'.join([str(x) for x in res[3]])))
