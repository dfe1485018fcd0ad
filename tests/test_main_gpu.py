"""The drop-in driver (gdmcf_b200/main.py, mirror of the reference's main.py) end to end on a tiny synthetic dataset,
and checkpoint / resume: 2 epochs + resume for 2 more == 4 epochs uninterrupted, bit for bit."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(tmp, name, extra):
    from gdmcf_b200 import main as M
    from gdmcf_b200.parse_args_util import parse_args
    argv = ["--synthetic", "600,500,15000", "--dims", "64", "--batch_size", "100", "--steps", "5", "--noise_scale", "0.01",
            "--topN", "[10, 20]", "--lr", "0.001", "--eval_every", "2", "--debug", "true", "--log_name", str(tmp),
            "--dataset", "tiny", "--out_name", name] + extra
    res = M.main(parse_args(argv))
    model = M.main.last_model
    return res, {k: v.detach().clone() for k, v in model.state_dict().items()}


def test_main_runs_and_resumes(tmp_path):
    res_full, sd_full = _run(tmp_path, "full", ["--epochs", "4"])
    assert res_full is not None and len(res_full) == 4 and all(len(r) == 2 for r in res_full)
    assert all(0.0 <= x <= 1.0 for r in res_full for x in r)
    _run(tmp_path, "part", ["--epochs", "2", "--checkpoint_every", "2"])
    ck = glob.glob(os.path.join(str(tmp_path), "tiny", "*", "part", "checkpoint.pt"))
    assert len(ck) == 1
    res_res, sd_res = _run(tmp_path, "resumed", ["--epochs", "4", "--resume", ck[0]])
    for k in sd_full:
        assert torch.equal(sd_full[k], sd_res[k]), k
    assert res_res == res_full
