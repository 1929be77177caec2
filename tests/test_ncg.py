"""Negative-control-gene statistics: K11 against the numpy oracle, and the selection flow."""
import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import ncg_oracle as NO
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api
from ruviii_b200.ncg import model_matrix, r_rank, supervisedFindNCG


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
        supervisedFindNCG(se, "a", ["b"], ["b"], apply_normalization=True, verbose=False)
    with pytest.raises(ValueError):                      # the reference stops on the missing expr.data.reg.bio (:164)
        supervisedFindNCG(se, "a", ["b"], ["b"], regress_out_uv_variables=True, regress_out_bio_variables=True, verbose=False)
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


def test_model_matrix_is_lm_model_matrix():
    cd = {"f": ["b", "a", "c", "a"], "x": [1.0, 2.0, 3.0, 5.0]}
    D = model_matrix(cd, ["f", "x"])                       # (Intercept) fb fc x -- treatment contrasts, levels sorted
    assert D.tolist() == [[1, 1, 0, 1.0], [1, 0, 0, 2.0], [1, 0, 1, 3.0], [1, 0, 0, 5.0]]
    expr = np.random.default_rng(0).normal(size=(7, 4))
    res = NO.lm_residuals(expr, cd, ["f"])                 # one factor: residuals = deviations from the level means
    assert np.allclose(res[:, [1, 3]].sum(axis=1), 0) and np.allclose(res[:, [0, 2]], 0)


@pytest.mark.gpu
def test_regress_out_matches_lm_residuals(engine):
    """K12 + K4 + K6 against numpy.linalg.lstsq: one chunk, several chunks (p > 32: a 40-level factor), a rank-deficient
    design (a factor nested in another: lm drops the aliased columns), and the statistics on the residuals (slot = -2)."""
    counts, cd = _design(m1=410, n=1203)
    expr = np.log2(counts + 1.0)
    rng = np.random.default_rng(3)
    cd["lane"] = ["L%02d" % i for i in rng.integers(0, 40, 410)]
    cd["site"] = ["S" + v[1] for v in cd["lane"]]                        # nested in lane: aliased
    for variables, p_rank in ((["plate", "libsize"], None), (["lane", "purity"], None), (["lane", "site", "libsize"], 41)):
        D = model_matrix(cd, variables)
        got = engine.regress_out(counts.T, D, apply_log=True, pseudo_count=1.0, return_residuals=True)
        ref = NO.lm_residuals(expr, cd, variables)
        assert got["rank"] == (p_rank or D.shape[1]) == np.linalg.matrix_rank(D)
        assert O.rel_frob(got["residuals"].T, ref) < 1e-11, variables
        # the residuals are orthogonal to every column of the model matrix
        assert np.max(np.abs(D.T @ got["residuals"])) < 1e-8 * np.abs(expr).max() * 410
    bio = np.asarray(cd["subtype"])
    F = engine.gene_stats(None, groups=np.unique(bio, return_inverse=True)[1].astype(np.int32), slot=-2)["F"]
    assert O.rel_frob(F, NO.row_oneway_equalvar_statistic(ref, bio)) < 1e-9
    rho = engine.gene_spearman(None, covariate=np.asarray(cd["purity"]), slot=-2)["rho"]
    assert np.max(np.abs(rho - NO.correls_spearman(cd["purity"], ref.T))) < 1e-10
    with pytest.raises(rv.RuviiiError):
        engine.regress_out(counts.T, np.ones((409, 1)))


@pytest.mark.gpu
def test_regress_out_on_resident_data(engine):
    from ruviii_b200.synth import make_problem
    p = make_problem(m1=260, n=1101, groups=20, rep=3, n_ctl=100, k=3, seed=78)
    groups = api.group_codes(p.sample_names + p.replicate_names)
    out = engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl, [3])
    cd = {"batch": ["b%d" % ((i // 5) % 8) for i in range(260)], "t": list(np.linspace(0, 1, 260) ** 2)}
    res = engine.regress_out(None, model_matrix(cd, ["batch", "t"]), slot=0, return_residuals=True)["residuals"]
    assert O.rel_frob(res.T, NO.lm_residuals(out["newY"][0].T, cd, ["batch", "t"])) < 1e-11
    res = engine.regress_out(None, model_matrix(cd, ["batch"]), slot=-1, return_residuals=True)["residuals"]
    assert O.rel_frob(res.T, NO.lm_residuals(p.Y.T, cd, ["batch"])) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["uv", "bio"])
def test_supervised_find_ncg_regress_out(engine, which):
    counts, cd = _design()
    se = rv.SummarizedExperiment(assays={"counts": counts}, col_names=["s%d" % i for i in range(counts.shape[1])], col_data=cd)
    api.set_engine(engine)
    try:
        se = supervisedFindNCG(se, "counts", ["subtype", "purity"], ["plate", "libsize"], no_ncg=150, apply_log=True, pseudo_count=1,
                               regress_out_uv_variables=which == "uv", regress_out_bio_variables=which == "bio", verbose=False)
    finally:
        api.set_engine(None)
    ref = NO.supervised_find_ncg(counts, cd, ["subtype", "purity"], ["plate", "libsize"], 150, True, 1.0,
                                 regress_out_uv=which == "uv", regress_out_bio=which == "bio")
    plain = NO.supervised_find_ncg(counts, cd, ["subtype", "purity"], ["plate", "libsize"], 150, True, 1.0)
    assert not np.array_equal(ref, plain)                                  # the option changes the selection
    # correlations are rounded to two decimals before ranking (:276-277): a residual that differs in the last bits can move
    # a gene across a rounding boundary, so allow a handful of swaps at the selection edge
    assert se.metadata["NCG"].sum() == ref.sum() and np.sum(se.metadata["NCG"] != ref) <= 4


class _StubEngine:
    """numpy stand-in for the engine's statistics entry points (tests only): lets the HOST flow of supervisedFindNCG -- which
    data every test reads, the level filter, rounding, ranking, the rank product -- run without a GPU."""

    def __init__(self):
        self.reg = None

    @staticmethod
    def _expr(X, apply_log, pseudo_count):
        X = np.asarray(X, dtype=np.float64)
        return np.log2(X + pseudo_count) if apply_log else X

    def regress_out(self, X, design, slot=-1, apply_log=False, pseudo_count=1.0, return_residuals=False):
        E = self._expr(X, apply_log, pseudo_count)                       # samples x genes
        self.reg = E - design @ np.linalg.lstsq(design, E, rcond=None)[0]
        return dict(rank=int(np.linalg.matrix_rank(design)), ms=0.0)

    def _data(self, X, slot, apply_log, pseudo_count):
        return self.reg if slot == -2 else self._expr(X, apply_log, pseudo_count)

    def gene_stats(self, X=None, groups=None, covariate=None, slot=-1, apply_log=False, pseudo_count=1.0):
        E = self._data(X, slot, apply_log, pseudo_count)
        out = dict(F=None, r=None, ms=0.0)
        if groups is not None:
            keep = np.asarray(groups) >= 0
            out["F"] = NO.row_oneway_equalvar_statistic(E[keep].T, np.asarray(groups)[keep])
        if covariate is not None:
            out["r"] = NO.correls_pearson(covariate, E)
        return out

    def gene_spearman(self, X=None, covariate=None, slot=-1):
        return dict(rho=NO.correls_spearman(covariate, self._data(X, slot, False, 1.0)), ms=0.0)


@pytest.mark.parametrize("opts", [dict(), dict(corr_method="spearman"), dict(regress_out_uv_variables=True),
                                  dict(regress_out_bio_variables=True)])
def test_ncg_host_flow_with_a_stub_engine(opts):
    counts, cd = _design(m1=120, n=301)
    se = rv.SummarizedExperiment(assays={"counts": counts}, col_names=["s%d" % i for i in range(counts.shape[1])], col_data=cd)
    api.set_engine(_StubEngine())
    try:
        got = supervisedFindNCG(se, "counts", ["subtype", "purity"], ["plate", "libsize"], no_ncg=40, apply_log=True, pseudo_count=1,
                                save_se_obj=False, verbose=False, **opts)
    finally:
        api.set_engine(None)
    ref = NO.supervised_find_ncg(counts, cd, ["subtype", "purity"], ["plate", "libsize"], 40, True, 1.0,
                                 corr_method=opts.get("corr_method", "pearson"), regress_out_uv=opts.get("regress_out_uv_variables", False),
                                 regress_out_bio=opts.get("regress_out_bio_variables", False))
    assert got.dtype == bool and np.array_equal(got, ref)


@pytest.mark.gpu
def test_engine_reproduces_the_base_r_side_fixtures(engine, golden_mid):
    """The CUDA statistics against the golden files replay.R checks with base R (lm, oneway.test, cor, svd): engine -> golden
    here, golden -> R in one command there."""
    g = golden_mid
    assay, batch, cov = g["assay"], np.asarray(g["meta"]["side_batch"]), g["side.covariate"][:, 0]
    codes = np.unique(batch, return_inverse=True)[1].astype(np.int32)
    st = engine.gene_stats(assay.T, groups=codes, covariate=cov)
    assert O.rel_frob(st["F"], g["side.anovaF"][:, 0]) < 1e-10 and np.max(np.abs(st["r"] - g["side.pearson"][:, 0])) < 1e-12
    assert np.max(np.abs(engine.gene_spearman(assay.T, covariate=cov)["rho"] - g["side.spearman"][:, 0])) < 1e-12
    res = engine.regress_out(assay.T, model_matrix({"batch": list(batch), "cov": list(cov)}, ["batch", "cov"]), return_residuals=True)
    assert res["rank"] == 6 and O.rel_frob(res["residuals"].T, g["side.lm_residuals"]) < 1e-11
    pc = engine.pca(assay.T, k=60, apply_log=False)                        # the full svd of computePCA (fast.pca = FALSE)
    assert np.allclose(pc["d"][:59], g["side.svd_d"][:59, 0], rtol=1e-9)
    assert O.rel_frob((pc["u"][:, :10] * pc["d"][:10]) @ pc["v"][:, :10].T, g["side.svd_rank10"]) < 1e-9
