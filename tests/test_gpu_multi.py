"""Multi-GPU parity (needs >= 2 B200s on one box; skipped otherwise): one process driving N gene shards
(`ruviii_create`, what an R session does) must reproduce the single-GPU result, and the one-process-per-GPU
launch of bench.py must run."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api
from ruviii_b200.synth import make_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _ngpu():
    return _cabi.load_library().ruviii_device_count()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_single_process_gene_shards(ndev):
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    p = make_problem(m1=700, n=9001, groups=90, rep=3, n_ctl=400, k=9, seed=21)      # odd n: ragged shards
    groups = api.group_codes(p.sample_names + p.replicate_names)
    prm = _cabi.make_params(want_ps_rows=True)
    with rv.Engine(devices=[0]) as e1:
        one = e1.normalise(prm, p.Y, p.replicate_data, groups, p.ctl, [4, 9], want_fullalpha=True, want_W=True, want_ps=True)
    with rv.Engine(devices=list(range(ndev))) as en:
        many = en.normalise(prm, p.Y, p.replicate_data, groups, p.ctl, [4, 9], want_fullalpha=True, want_W=True, want_ps=True)
        assert many["info"].n_shards == ndev
    ref = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 9, apply_log=False)
    for res in (one, many):
        assert O.rel_frob(np.vstack([res["newY"][1], res["newY_ps"][1]]), ref["newY"]) < 1e-9
    # summation order over shards differs from the single-GPU order only at rounding level
    assert O.rel_frob(many["newY"][0], one["newY"][0]) < 1e-11
    assert O.rel_frob(many["W"][1], one["W"][1]) < 1e-8
    assert O.rel_frob(O.align_rows_sign(many["fullalpha"], one["fullalpha"]), one["fullalpha"]) < 1e-8


def test_torchrun_two_ranks_bench():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "1",
           "--workload", "tiny"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["gpu_launches"] > 0
