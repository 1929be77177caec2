"""CPU tests of the host-side mirror of the R interface: argument handling and error behaviour that the
reference implements before any arithmetic (ruvIII.R:62-63,80-84,104-114; fastruvIII.R:65-68)."""
import numpy as np
import pytest

import ruviii_b200 as rv
from ruviii_b200 import api
from ruviii_b200.synth import make_problem


@pytest.fixture(scope="module")
def small():
    p = make_problem(m1=30, n=50, groups=4, rep=2, n_ctl=10, k=2, seed=3, per_cell=3)
    se = rv.SummarizedExperiment(assays={"logcounts": p.assay}, col_names=p.sample_names)
    return p, se, rv.ReplicateData(p.replicate_data, p.replicate_names)


def test_k_null_stops_like_r(small):
    p, se, rd = small
    for fn in (rv.ruvIII, rv.fastruvIII, rv.ruvIIIMultipleK):
        with pytest.raises(ValueError, match="k cannot be 0"):
            fn(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=None, verbose=False)


def test_fast_rejects_singleton_prps_sets(small):
    p, se, rd = small
    bad = rv.ReplicateData(p.replicate_data, ["u%d" % i for i in range(len(p.replicate_names))])
    with pytest.raises(ValueError, match="single sample"):
        rv.fastruvIII(se, "logcounts", apply_log=False, replicate_data=bad, ctl=p.ctl, k=1, verbose=False)


def test_group_codes_and_tological():
    codes = api.group_codes(["a", "b", "a", "c", "b"])
    assert codes.tolist() == [0, 1, 0, 2, 1] and codes.dtype == np.int32
    assert api._tological(np.array([0, 3]), 5).tolist() == [True, False, False, True, False]
    with pytest.raises(ValueError):
        api._tological(np.array([True, False]), 5)


def test_shape_validation_happens_before_the_gpu(small):
    p, se, rd = small
    with pytest.raises(KeyError):
        rv.ruvIII(se, "nope", replicate_data=rd, ctl=p.ctl, k=1, verbose=False)
    with pytest.raises(ValueError, match="one column per gene"):
        rv.ruvIII(se, "logcounts", replicate_data=rv.ReplicateData(p.replicate_data[:, :-1], p.replicate_names),
                  ctl=p.ctl, k=1, verbose=False)
    with pytest.raises(ValueError, match="one row per gene"):
        rv.ruvIII(se, "logcounts", replicate_data=rd, ctl=p.ctl, k=1, eta=np.ones((2, 49)), verbose=False)


def test_eta_design_matrix_follows_ruv1():
    """ruv::RUV1 (called at ruvIII.R:116): a single number means intercept only; a gene-wise matrix (n x q, R's orientation, or
    q x n) gets a column of ones when include.intercept and no constant column is already there (design.matrix)."""
    n = 7
    assert api._design_matrix_t(1, n, True).tolist() == [[1.0] * n]
    x = np.arange(n, dtype=float)
    e = api._design_matrix_t(x[:, None], n, True)                  # n x 1, R's orientation
    assert e.shape == (2, n) and e[0].tolist() == [1.0] * n and np.array_equal(e[1], x)
    assert api._design_matrix_t(x[None, :], n, False).shape == (1, n)
    both = np.vstack([np.ones(n), x])
    assert api._design_matrix_t(both, n, True).shape == (2, n)     # already has its intercept


def test_no_gpu_means_error_not_fallback(small):
    from ruviii_b200 import _cabi
    if _cabi.load_library().ruviii_device_count() > 0:
        pytest.skip("GPU present")
    p, se, rd = small
    api.set_engine(None)
    with pytest.raises(rv.RuviiiError, match="no CPU fallback"):
        rv.ruvIII(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=1, verbose=False)
