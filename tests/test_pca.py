"""computePCA fast path against the numpy oracle.  Singular vectors are sign-free: compared after
aligning each component's sign; tolerance 1e-9 relative Frobenius like the rest of the path."""
import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import pca_oracle as PCAO
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api
from ruviii_b200.pca import computePCA
from ruviii_b200.synth import make_problem


def test_argument_errors_match_the_reference():
    se = rv.SummarizedExperiment(assays={"a": np.ones((4, 3))}, col_names=["x", "y", "z"])
    with pytest.raises(ValueError, match="number of PCs"):
        computePCA(se, nb_pcs=0, verbose=False)
    with pytest.raises(ValueError, match="at least an assay name"):
        computePCA(se, assay_names=None, verbose=False)


def _check(res, ref, k):
    assert np.allclose(res["d"], ref["d"], rtol=1e-10, atol=0)
    u = O.align_rows_sign(res["u"].T, ref["u"].T).T
    v = O.align_rows_sign(res["v"].T, ref["v"].T).T
    assert O.rel_frob(u, ref["u"]) < 1e-8 and O.rel_frob(v, ref["v"]) < 1e-8
    # sign-free: the rank-k reconstruction
    assert O.rel_frob((res["u"] * res["d"]) @ res["v"].T, (ref["u"] * ref["d"]) @ ref["v"].T) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("shape,k,scale", [((300, 1203), 10, False), ((200, 515), 4, True), ((150, 90), 20, False)])
def test_pca_matches_oracle(engine, shape, k, scale):
    """k = 20 > 16 and the 150-sample case exercise the cuSOLVER route; the others the block subspace iteration."""
    m1, n = shape
    p = make_problem(m1=m1, n=n, groups=10, rep=3, n_ctl=40, k=3, seed=m1 + n, counts=True)
    res = engine.pca(p.assay.T, k=k, apply_log=True, pseudo_count=1.0, center=True, scale=scale)
    ref = PCAO.fast_pca(p.assay, k, True, 1.0, True, scale)
    _check(res, ref, k)
    assert res["u"].shape == (m1, k) and res["v"].shape == (n, k)


@pytest.mark.gpu
def test_pca_on_resident_outputs_and_r_api(engine):
    """PCA of the normalised assays without leaving the device (slot = output block), and the computePCA mirror."""
    p = make_problem(m1=260, n=1101, groups=20, rep=3, n_ctl=100, k=4, seed=77)
    groups = api.group_codes(p.sample_names + p.replicate_names)
    out = engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl, [2, 4])
    for slot in (0, 1):
        res = engine.pca(None, k=6, slot=slot)
        ref = PCAO.fast_pca(out["newY"][slot].T, 6, apply_log=False)
        _check(res, ref, 6)
    raw = engine.pca(None, k=6, slot=-1)                                   # the uploaded assay itself
    _check(raw, PCAO.fast_pca(p.assay, 6, apply_log=False), 6)
    again = engine.download()                                              # outputs are untouched by the PCA
    assert np.array_equal(again["newY"][1], out["newY"][1])
    se = rv.SummarizedExperiment(assays={"logcounts": p.assay, "RUV_K4_on_logcounts": out["newY"][1].T}, col_names=p.sample_names)
    api.set_engine(engine)
    try:
        se = computePCA(se, "All", apply_log=False, nb_pcs=5, verbose=False)
    finally:
        api.set_engine(None)
    for name in ("logcounts", "RUV_K4_on_logcounts"):
        got = se.metadata["metric"][name]["fastPCA"]
        ref = PCAO.fast_pca(se.assays[name], 5, apply_log=False)
        assert np.array_equal(got["variation"], ref["variation"])
        _check(got["sing_val"], ref, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,k", [((150, 400), 40), ((150, 400), 97), ((120, 50), 50)])
def test_pca_beyond_32_components(engine, shape, k):
    """k > 32: library eigensolver for the k leading pairs, projection in chunks of 32 (one full chunk + a ragged one; k = n)."""
    m1, n = shape
    p = make_problem(m1=m1, n=n, groups=10, rep=3, n_ctl=20, k=3, seed=m1 + n + k)
    res = engine.pca(p.assay.T, k=k, apply_log=False, center=True, scale=False)
    ref = PCAO.fast_pca(p.assay, k, apply_log=False)
    assert res["u"].shape == (m1, k) and res["v"].shape == (n, k)
    assert np.allclose(res["d"], ref["d"], rtol=1e-9, atol=1e-9 * ref["d"][0])
    assert O.rel_frob((res["u"] * res["d"]) @ res["v"].T, (ref["u"] * ref["d"]) @ ref["v"].T) < 1e-9
    lead = 20                                                    # well-separated components: vectors agree one by one
    u = O.align_rows_sign(res["u"][:, :lead].T, ref["u"][:, :lead].T).T
    assert O.rel_frob(u, ref["u"][:, :lead]) < 1e-8
    assert np.max(np.abs(res["u"].T @ res["u"] - np.eye(k))) < 1e-9                      # orthonormal columns throughout
    with pytest.raises(rv.RuviiiError):
        engine.pca(p.assay.T, k=min(m1, n) + 1)


@pytest.mark.gpu
def test_compute_pca_full_svd(engine):
    """fast.pca = FALSE (computePCA.R:150-190): every singular triplet, saved under metadata$metric$<assay>$PCA.  The centred
    matrix has rank m1 - 1: the last value is rounding noise in both implementations and is left out of the comparison."""
    p = make_problem(m1=70, n=300, groups=6, rep=3, n_ctl=20, k=3, seed=5, counts=True)
    se = rv.SummarizedExperiment(assays={"counts": p.assay}, col_names=p.sample_names)
    api.set_engine(engine)
    try:
        se = computePCA(se, "counts", apply_log=True, fast_pca=False, nb_pcs=None, verbose=False)
    finally:
        api.set_engine(None)
    got = se.metadata["metric"]["counts"]["PCA"]
    ref = PCAO.fast_pca(p.assay, 70, apply_log=True)
    d, dr = got["sing_val"]["d"], ref["d"]
    assert d.shape == (70,) and got["sing_val"]["u"].shape == (70, 70) and got["sing_val"]["v"].shape == (300, 70)
    assert np.allclose(d[:69], dr[:69], rtol=1e-9) and d[69] < 1e-6 * d[0]
    assert np.array_equal(got["variation"], ref["variation"])
    sv = got["sing_val"]
    assert O.rel_frob((sv["u"][:, :69] * d[:69]) @ sv["v"][:, :69].T, (ref["u"][:, :69] * dr[:69]) @ ref["v"][:, :69].T) < 1e-9


class _StubPcaEngine:
    """numpy stand-in for ruviii_pca (tests only): the HOST flow of computePCA -- which k is asked for, where the result is
    saved, the percentages -- without a GPU."""

    def __init__(self):
        self.calls = []

    def pca(self, X, k=10, apply_log=False, pseudo_count=1.0, center=True, scale=False, **_):
        self.calls.append(k)
        ref = PCAO.fast_pca(np.asarray(X).T, k, apply_log, pseudo_count, center, scale)
        return dict(u=ref["u"], d=ref["d"], v=ref["v"])


def test_compute_pca_host_flow_with_a_stub_engine():
    assay = np.random.default_rng(9).normal(8.0, 2.0, (90, 40))           # genes x samples
    se = rv.SummarizedExperiment(assays={"a": assay, "b": assay * 2.0}, col_names=["s%d" % i for i in range(40)])
    stub = _StubPcaEngine()
    api.set_engine(stub)
    try:
        se = computePCA(se, "All", apply_log=False, fast_pca=True, nb_pcs=7, verbose=False)
        se = computePCA(se, "b", apply_log=False, fast_pca=False, nb_pcs=None, verbose=False)
    finally:
        api.set_engine(None)
    assert stub.calls == [7, 7, 40]                                         # svd() keeps min(samples, genes) triplets
    assert set(se.metadata["metric"]["a"]) == {"fastPCA"} and set(se.metadata["metric"]["b"]) == {"fastPCA", "PCA"}
    full = se.metadata["metric"]["b"]["PCA"]
    assert full["sing_val"]["u"].shape == (40, 40) and full["sing_val"]["v"].shape == (90, 40)
    assert abs(full["variation"].sum() - 100.0) < 2.0 and full["variation"][0] >= full["variation"][1]
