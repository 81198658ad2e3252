"""The reference's own launcher and checkpoints against the drop-in modules.

These tests execute the UNMODIFIED reference from /root/reference (imported in a subprocess, never copied), so they run in
the build container only -- the GPU box has no /root/reference and the CUDA path is covered there by tests/test_gpu_*.py.
The engine runs on tests/emulate_ops.py in fp32: what is under test is the integration surface (generate_model, the DDP
wrap, autograd through the drop-in forward, optim.SGD on the aliased parameters, checkpoint formats), which is host code.
"""
import json
import os
import subprocess
import sys

import pytest

REF = os.environ.get("CSTP_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main_byol.py")),
                                reason="the reference checkout is only present in the build container")


def _run(script, out):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", script), str(out)], capture_output=True, text=True,
                       timeout=1500, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return json.load(open(os.path.join(out, "result.json")))


def test_unmodified_launcher_trains_on_the_dropin_backend(tmp_path):
    """main_byol.py (main_worker + train_BYOL, argparse options, DistributedSampler loader, DDP-wrapped model from
    generate_model, CrossEntropyLoss x 6, --loss_weight, clip_grad_norm_, optim.SGD, cosine schedule, TSV Logger) for two
    epochs: every logged column equals the oracle's value for the same two optimiser steps."""
    res = _run("run_reference_launcher.py", tmp_path)
    assert res["wrapper"] == "DistributedDataParallel"                 # generate_model's DDP wrap (models/model.py:97-103)
    assert res["model_class"] == "cstp_b200.models.pace.r21d_byol.R21DBYOL"
    rows = res["rows"]
    assert rows[0] == ["epoch", "loss", "loss_byol", "loss_pred_spa", "loss_pred_tem", "loss_pred_pb", "loss_pred_rot", "acc", "lr"]
    assert [r[0] for r in rows[1:]] == ["1", "2"]
    assert rows[1][8] == "1e-05" and rows[2][8] == "0.03"              # epoch 1 runs at min_lr (main_byol.py:252-258)
    for row, ref in zip(rows[1:], res["oracle"]):
        got = [float(v) for v in row[1:7]]
        want = [ref["loss"], ref["loss_byol"], ref["ce"][0], ref["ce"][1], (ref["ce"][2] + ref["ce"][3]) / 2,
                (ref["ce"][4] + ref["ce"][5]) / 2]                         # main_byol.py:81-84: pb / rot columns are means
        assert all(abs(a - b) < 1e-5 * abs(b) for a, b in zip(got, want)), (got, want)
        assert row[7] == ""                                             # "acc": None
    assert res["weight_rel"] < 1e-3                                     # the weights after two SGD steps follow the oracle's


def test_checkpoints_written_by_the_reference_and_by_us_are_interchangeable(tmp_path):
    res = _run("run_reference_checkpoint.py", tmp_path)
    assert res["resume_epoch"] == 100                                   # main_byol.py:214-215: parsed from the file name
    assert res["weights_equal_after_load"] and res["momentum_equal_after_load"]
    for a, b in ((res["ours_step2"], res["ref_step2"]), (res["ours_step3"], res["ref2_step3"])):
        assert abs(a[0] - b[0]) < 1e-5 * b[0] and abs(a[1] - b[1]) < 1e-5 * b[1], (a, b)
    assert res["weights_rel_after_step2"] < 2e-2 and res["weights_rel_after_step3"] < 2e-2
    missing, unexpected, epoch, arch = res["ref_loads_ours"]
    assert missing == [] and unexpected == [] and epoch == 102 and arch == "r21d_byol-1"
    assert res["ft_backbone_equal"]                                     # neq_load_customized: online_net.* carried over
    assert res["ft_head_keys"] == ["classify.bias", "classify.weight", "cls_bn.bias", "cls_bn.num_batches_tracked",
                                   "cls_bn.running_mean", "cls_bn.running_var", "cls_bn.weight"]
