"""Property tests (hypothesis) of the host-side bookkeeping that feeds the device kernels: ranks, factor codes, PRPS index
sets and the dense-design generator's slicing.  CPU only."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from scipy.stats import rankdata

from ruviii_b200 import api, prps
from ruviii_b200.ncg import r_rank
from ruviii_b200.sharding import gene_range
from ruviii_b200.synth import make_dense_design


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(-5, 5), min_size=1, max_size=60))
def test_r_rank_is_average_rank(xs):
    x = np.asarray(xs, dtype=np.float64)
    assert np.array_equal(r_rank(x), rankdata(x, method="average"))


@settings(max_examples=40, deadline=None)
@given(st.lists(st.sampled_from(["a", "b", "c", "d", "e_1", "e-1"]), min_size=1, max_size=40))
def test_group_codes_are_factor_codes(names):
    codes = api.group_codes(names)
    levels = sorted(set(names))
    assert [levels[c] for c in codes] == names            # as.integer(factor(names)) - 1


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 6), st.integers(2, 6), st.integers(20, 120), st.integers(2, 4), st.integers(0, 2 ** 31 - 1))
def test_categorical_prps_sets_are_disjoint_cells(n_bio, n_batch, m1, min_s, seed):
    rng = np.random.default_rng(seed)
    bio = ["b%d" % i for i in rng.integers(0, n_bio, m1)]
    uv = ["u_%d" % i for i in rng.integers(0, n_batch, m1)]
    try:
        sets, names = prps.prps_sets_categorical(bio, uv, "plate", min_s)
    except ValueError:
        return                                              # not enough samples anywhere: the reference stops too
    flat = np.concatenate(sets)
    assert len(flat) == len(set(flat.tolist()))            # a sample sits in one (biology, batch) cell
    for idx, nm in zip(sets, names):
        assert len(idx) >= min_s
        assert len({bio[i] for i in idx}) == 1 and len({uv[i] for i in idx}) == 1
        assert nm == "%s||plate" % bio[idx[0]]
    from collections import Counter
    assert min(Counter(names).values()) >= 2               # bio.dist > 1: at least two pseudo-samples per set


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 8), st.integers(30, 400))
def test_dense_design_slices_tile_the_full_matrix(world, n):
    Y, gid, ctl = make_dense_design(m=21, n=n, n_ctl=min(10, n), block=64)
    parts = [make_dense_design(m=21, n=n, n_ctl=min(10, n), block=64, cols=gene_range(n, r, world)) for r in range(world)]
    assert np.array_equal(np.concatenate([p[0] for p in parts], axis=1), Y)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), ctl)
