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


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_gene_side_gram_over_shards(ndev):
    """BASELINE config 5's regime (more replicate rows than genes) on several GPUs: t(Z) Z couples the gene shards, so
    the engine re-splits the contraction over the rows of Z with an all-to-all and allreduces the n x n partials."""
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    from ruviii_b200.synth import make_dense_design
    Y, gid, ctl = make_dense_design(m=1203, n=389, group_size=3, n_ctl=120, k=6, seed=31)     # ragged shards and row blocks
    PS = np.zeros((0, 389))
    with rv.Engine(devices=[0]) as e1:
        one = e1.normalise(_cabi.make_params(), Y, PS, gid, ctl, [3, 6], want_fullalpha=True, want_W=True)
    with rv.Engine(devices=list(range(ndev))) as en:
        many = en.normalise(_cabi.make_params(), Y, PS, gid, ctl, [3, 6], want_fullalpha=True, want_W=True)
        assert many["info"].n_shards == ndev and many["info"].gram_side == 2 and many["info"].gram_order == 389
    ref = O.ruvIII_literal(Y.T, ["g%d" % g for g in gid], PS, [], ctl, 6, apply_log=False)
    for res in (one, many):
        assert O.rel_frob(res["newY"][1], ref["newY"]) < 1e-9
    assert O.rel_frob(many["newY"][0], one["newY"][0]) < 1e-10
    assert O.rel_frob(O.align_rows_sign(many["fullalpha"], one["fullalpha"]), one["fullalpha"]) < 1e-8


def test_aux_entry_points_over_two_shards():
    """The 8f entry points (PRPS means, gene statistics, fast PCA) shard the genes like the path itself: two shards must
    reproduce one (PCA: the Gram partials are allreduced; statistics and means are per gene)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    p = make_problem(m1=300, n=1501, groups=20, rep=3, n_ctl=100, k=3, seed=91)                 # odd n: ragged shards
    rng = np.random.default_rng(3)
    sets = [np.sort(rng.choice(300, 3, replace=False)) for _ in range(40)]
    batch = (np.arange(300) % 7).astype(np.int32)
    cov = rng.gamma(4.0, 1.0, 300)
    with rv.Engine(devices=[0]) as e1:
        ps1 = e1.prps(p.Y, sets)
        gs1 = e1.gene_stats(p.Y, groups=batch, covariate=cov)
        pc1 = e1.pca(p.Y, k=5)
    with rv.Engine(devices=[0, 1]) as e2:
        ps2 = e2.prps(p.Y, sets)
        gs2 = e2.gene_stats(p.Y, groups=batch, covariate=cov)
        pc2 = e2.pca(p.Y, k=5)
    assert np.array_equal(ps1, ps2)                                     # per-gene arithmetic: identical bits
    assert np.array_equal(gs1["F"], gs2["F"]) and np.array_equal(gs1["r"], gs2["r"])
    assert np.allclose(pc1["d"], pc2["d"], rtol=1e-12, atol=0)
    assert O.rel_frob(O.align_rows_sign(pc2["u"].T, pc1["u"].T), pc1["u"].T) < 1e-9
    assert O.rel_frob(O.align_rows_sign(pc2["v"].T, pc1["v"].T), pc1["v"].T) < 1e-9
    assert O.rel_frob(ps1, np.stack([p.Y[s].mean(axis=0) for s in sets])) < 1e-14


def test_torchrun_two_ranks_gene_side(tmp_path):
    """One process per GPU: the per-rank gene counts travel through the plan-time allreduce."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    code = (
        "import os, numpy as np, torch.distributed as dist\n"
        "from ruviii_b200 import _cabi\n"
        "from ruviii_b200.sharding import engine_from_torch_distributed, gene_range\n"
        "from ruviii_b200.synth import make_dense_design\n"
        "from oracle import ruviii_oracle as O\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "Y, gid, ctl = make_dense_design(m=903, n=301, group_size=3, n_ctl=90, k=5, seed=33)\n"
        "g0, g1 = gene_range(301, r, w)\n"
        "eng = engine_from_torch_distributed()\n"
        "res = eng.normalise(_cabi.make_params(), np.ascontiguousarray(Y[:, g0:g1]), np.zeros((0, g1 - g0)), gid, ctl[g0:g1], [5])\n"
        "assert res['info'].gram_side == 2 and res['info'].gram_order == 301, (res['info'].gram_side, res['info'].gram_order)\n"
        "ref = O.ruvIII_literal(Y.T, ['g%d' % g for g in gid], np.zeros((0, 301)), [], ctl, 5, apply_log=False)\n"
        "err = O.rel_frob(res['newY'][0], ref['newY'][:, g0:g1])\n"
        "print('rank', r, 'err', err)\n"
        "assert err < 1e-9\n"
        "eng.close(); dist.destroy_process_group()\n")
    script = tmp_path / "gene_side_rank.py"
    script.write_text(code)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", str(script)]
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-2500:])


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
