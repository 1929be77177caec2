"""Negative-control-gene statistics: K11 against the numpy oracle, and the selection flow."""
import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import ncg_oracle as NO
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api
from ruviii_b200.ncg import r_rank, supervisedFindNCG


def _design(m1=330, n=1203, seed=4):
    rng = np.random.default_rng(seed)
    bio = np.array(["basal", "lumA", "lumB", "her2", "normal"])[rng.integers(0, 5, m1)]
    batch = np.array(["p%02d" % i for i in rng.integers(0, 9, m1)])
    batch[:2] = "tiny"                                            # a level with two samples: dropped from the UV ANOVA
    lib = rng.gamma(6.0, 1.0, m1)
    purity = rng.uniform(0.2, 1.0, m1)
    be = rng.normal(0, 0.6, (n, 5))[:, np.searchsorted(np.unique(bio), bio)] * (rng.random(n) < 0.3)[:, None]
    bt = rng.normal(0, 0.5, (n, 10))[:, np.searchsorted(np.unique(batch), batch)]
    mu = rng.normal(7, 2, n)[:, None]
    logc = mu + be + bt + 0.4 * rng.normal(size=(n, 1)) * np.log(lib)[None, :] + rng.normal(0, 0.3, (n, m1))
    counts = np.round(np.exp2(np.clip(logc, 0, 18)))
    col_data = {"subtype": list(bio), "plate": list(batch), "libsize": list(lib), "purity": list(purity)}
    return counts, col_data


def test_rank_matches_r_semantics():
    x = np.array([3.0, 1.0, 3.0, 2.0, 1.0, 1.0])
    assert r_rank(x).tolist() == [5.5, 2.0, 5.5, 4.0, 2.0, 2.0]            # rank(c(3,1,3,2,1,1))
    assert np.array_equal(r_rank(x), NO.rank_average(x))


def test_unbuilt_options_are_loud():
    se = rv.SummarizedExperiment(assays={"a": np.ones((4, 3))}, col_names=["x", "y", "z"], col_data={"b": ["u", "v", "u"]})
    with pytest.raises(NotImplementedError):
        supervisedFindNCG(se, "a", ["b"], ["b"], regress_out_uv_variables=True, verbose=False)
    with pytest.raises(NotImplementedError):
        supervisedFindNCG(se, "a", ["b"], ["b"], corr_method="kendall", verbose=False)


@pytest.mark.gpu
def test_gene_stats_kernel_matches_oracle(engine):
    counts, cd = _design()
    expr = np.log2(counts + 1.0)
    g = np.asarray(cd["plate"])
    codes = np.unique(g, return_inverse=True)[1].astype(np.int32)
    lev, cnt = np.unique(g, return_counts=True)
    codes[np.isin(g, lev[cnt < 3])] = -1
    res = engine.gene_stats(counts.T, groups=codes, covariate=np.asarray(cd["libsize"]), apply_log=True, pseudo_count=1.0)
    keep = codes >= 0
    assert O.rel_frob(res["F"], NO.row_oneway_equalvar_statistic(expr[:, keep], g[keep])) < 1e-10
    assert np.max(np.abs(res["r"] - NO.correls_pearson(np.asarray(cd["libsize"])[keep], expr[:, keep].T))) < 1e-12
    # few large groups (5 biologies): every group is cut into pieces, the result may not depend on the cut
    bio = np.asarray(cd["subtype"])
    F2 = engine.gene_stats(expr.T, groups=np.unique(bio, return_inverse=True)[1])["F"]
    assert O.rel_frob(F2, NO.row_oneway_equalvar_statistic(expr, bio)) < 1e-10
    r2 = engine.gene_stats(expr.T, covariate=np.asarray(cd["purity"]))["r"]
    assert np.max(np.abs(r2 - NO.correls_pearson(cd["purity"], expr.T))) < 1e-12
    with pytest.raises(rv.RuviiiError):
        engine.gene_stats(expr.T, groups=np.zeros(expr.shape[1], dtype=np.int32))      # one group: no ANOVA


@pytest.mark.gpu
def test_gene_stats_on_resident_data(engine):
    from ruviii_b200.synth import make_problem
    p = make_problem(m1=260, n=1101, groups=20, rep=3, n_ctl=100, k=3, seed=78)
    groups = api.group_codes(p.sample_names + p.replicate_names)
    out = engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl, [3])
    batch = (np.arange(260) // 5) % 8
    F = engine.gene_stats(None, groups=batch.astype(np.int32), slot=0)["F"]
    assert O.rel_frob(F, NO.row_oneway_equalvar_statistic(out["newY"][0].T, batch)) < 1e-10


@pytest.mark.gpu
def test_supervised_find_ncg_selects_the_oracle_genes(engine):
    counts, cd = _design()
    se = rv.SummarizedExperiment(assays={"counts": counts}, col_names=["s%d" % i for i in range(counts.shape[1])], col_data=cd)
    api.set_engine(engine)
    try:
        se = supervisedFindNCG(se, "counts", ["subtype", "purity"], ["plate", "libsize"], no_ncg=150, apply_log=True,
                               pseudo_count=1, verbose=False)
    finally:
        api.set_engine(None)
    got = se.metadata["NCG"]
    ref = NO.supervised_find_ncg(counts, cd, ["subtype", "purity"], ["plate", "libsize"], 150, True, 1.0)
    assert got.dtype == bool and got.sum() == ref.sum()
    assert np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("m1", [330, 1500, 2049])
def test_spearman_kernel_matches_oracle(engine, m1):
    """K11b: shared-memory sort + tie-averaged ranks + Pearson of the ranks, against scipy's rankdata.  Counts have long runs
    of ties (zeros, small integers); m1 = 2049 crosses a power of two (padding keys)."""
    counts, cd = _design(m1=m1, n=611)
    counts[:40] = np.minimum(counts[:40], 3.0)                     # genes that are almost all ties
    counts[5] = 0.0                                                 # a constant gene: rho is 0 / 0
    lib = np.asarray(cd["libsize"])
    lib[::7] = lib[0]                                               # ties in the covariate too
    res = engine.gene_spearman(counts.T, covariate=lib)
    ref = NO.correls_spearman(lib, counts.T)
    ok = np.isfinite(ref)
    assert ok.sum() >= 500 and not np.isfinite(res["rho"][5])          # some of the clipped genes are constant too
    assert np.max(np.abs(res["rho"][ok] - ref[ok])) < 1e-12
    # ranks do not change under log2(x + 1)
    res2 = engine.gene_spearman(np.log2(counts.T + 1.0), covariate=lib)
    assert np.max(np.abs(res2["rho"][ok] - ref[ok])) < 1e-12


@pytest.mark.gpu
def test_supervised_find_ncg_spearman(engine):
    counts, cd = _design()
    se = rv.SummarizedExperiment(assays={"counts": counts}, col_names=["s%d" % i for i in range(counts.shape[1])], col_data=cd)
    api.set_engine(engine)
    try:
        se = supervisedFindNCG(se, "counts", ["subtype", "purity"], ["plate", "libsize"], no_ncg=150, apply_log=True,
                               pseudo_count=1, corr_method="spearman", verbose=False)
    finally:
        api.set_engine(None)
    ref = NO.supervised_find_ncg(counts, cd, ["subtype", "purity"], ["plate", "libsize"], 150, True, 1.0, corr_method="spearman")
    assert np.array_equal(se.metadata["NCG"], ref)
