"""CPU tests of the oracle itself: literal vs structural restatement, golden fixtures, and the algebraic
identities the engine relies on.  The reference has no tests of its own, so
these are what pins the checker (PARITY UNPINNED against R, see oracle/ruviii_oracle.py)."""
import numpy as np
import pytest

from oracle import ruviii_oracle as O
from ruviii_b200.synth import make_dense_design, make_problem

TOL = 1e-11


@pytest.fixture(scope="module")
def prob():
    return make_problem(m1=150, n=700, groups=11, rep=3, n_ctl=80, k=5, seed=7)


def _lit(p, k, **kw):
    return O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, k,
                            apply_log=False, **kw)


def test_golden_tiny_matches_literal(golden_tiny):
    g = golden_tiny
    m = g["meta"]
    L = O.ruvIII_literal(g["assay"], m["sample_names"], g["replicate_data"], m["replicate_names"], g["ctl"], m["k"],
                         apply_log=False)
    assert O.rel_frob(L["newY"], g["ruvIII.newY"]) < 1e-13
    assert L["fullalpha"].shape == tuple(m["dims"]["ruvIII.fullalpha"]) == (2, 6)   # r = min(m - p, sum(ctl)) = 2
    # hand check of residop on the tiny case: group A = PS rows 0,1 -> residual of row 0 is (row0 - row1) / 2
    M = O.replicate_matrix(m["sample_names"] + m["replicate_names"])
    Y = np.vstack([g["assay"].T, g["replicate_data"]])
    Y0 = O.residop(Y, M)
    assert np.allclose(Y0[:4], 0.0, atol=1e-14)                     # singleton rows vanish (appendix A-1)
    assert np.allclose(Y0[4], (Y[4] - Y[5]) / 2.0, atol=1e-14)
    assert np.allclose(Y0[4] + Y0[5], 0.0, atol=1e-14)


def test_golden_mid_matches_both_flavours(golden_mid):
    g = golden_mid
    m = g["meta"]
    S = O.ruvIII_structural(g["assay"], m["sample_names"], g["replicate_data"], m["replicate_names"], g["ctl"],
                            m["k"], apply_log=False)
    assert O.rel_frob(S["newY"], g["ruvIII.newY"]) < TOL
    assert O.rel_frob(S["W"] @ S["fullalpha"][:m["k"]], g["ruvIII.W"] @ g["ruvIII.fullalpha"][:m["k"]]) < 1e-10
    F = O.fastruvIII_structural(g["assay"], g["replicate_data"], m["replicate_names"], g["ctl"], m["k"],
                                apply_log=False)
    assert O.rel_frob(F["newY"], g["fastruvIII.newY"]) < TOL
    fa = O.align_rows_sign(F["fullalpha"], g["fastruvIII.fullalpha"])
    assert O.rel_frob(fa, g["fastruvIII.fullalpha"]) < 1e-9
    MK = O.ruvIIIMultipleK_structural(g["assay"], m["sample_names"], g["replicate_data"], m["replicate_names"],
                                      g["ctl"], m["ks_multi"], apply_log=False)
    for k in m["ks_multi"]:
        assert O.rel_frob(MK[k]["newY"], g["ruvIIIMultipleK.k%d.newY" % k]) < TOL
    Lc = O.ruvIII_structural(g["counts"], m["sample_names"], g["replicate_data"], m["replicate_names"], g["ctl"],
                             m["k"], apply_log=True, pseudo_count=1.0)
    assert O.rel_frob(Lc["newY"], g["ruvIII.log.newY"]) < TOL


@pytest.mark.parametrize("k", [1, 3, 5])
def test_literal_vs_structural(prob, k):
    L = _lit(prob, k)
    S = O.ruvIII_structural(prob.assay, prob.sample_names, prob.replicate_data, prob.replicate_names, prob.ctl, k,
                            apply_log=False)
    assert O.rel_frob(S["newY"], L["newY"]) < TOL
    assert L["fullalpha"].shape[0] == min(33 - 11, int(prob.ctl.sum()))          # r = min(m - p, sum(ctl))
    fa = O.align_rows_sign(S["fullalpha"][:k], L["fullalpha"][:k])
    assert O.rel_frob(fa, L["fullalpha"][:k]) < 1e-9
    assert O.rel_frob(S["W"] @ S["fullalpha"][:k], L["W"] @ L["fullalpha"][:k]) < 1e-10


def test_fast_literal_vs_structural_and_offset(prob):
    LF = O.fastruvIII_literal(prob.assay, prob.replicate_data, prob.replicate_names, prob.ctl, 4, apply_log=False)
    SF = O.fastruvIII_structural(prob.assay, prob.replicate_data, prob.replicate_names, prob.ctl, 4, apply_log=False)
    assert O.rel_frob(SF["newY"], LF["newY"]) < TOL
    # the two entry points differ by a row-constant offset (mu over m rows vs m1 rows, appendix A-6)
    L = _lit(prob, 4)
    d = LF["newY"] - L["newY"][:L["m1"]]
    assert np.linalg.norm(d - d.mean(axis=0, keepdims=True)) < 1e-9 * np.linalg.norm(L["newY"])
    assert np.linalg.norm(d) > 1e-6 * np.linalg.norm(L["newY"])


def test_residop_is_group_mean_subtraction(prob):
    Y = np.vstack([prob.Y, prob.replicate_data])
    names = prob.sample_names + prob.replicate_names
    Y0 = O.residop(Y, O.replicate_matrix(names))
    gid = O.group_ids_from_names(names)
    active, Y0a, p = O._group_mean_residual(Y, gid)
    assert p == 150 + 11
    assert np.abs(Y0[:150]).max() < 1e-12
    assert O.rel_frob(Y0a, Y0[active]) < 1e-13


def test_subspace_invariance_and_multik_nesting(prob):
    """newY depends only on span(alpha) (appendix A-4); W_k alpha_k is a prefix sum of rank-1 terms (A-7)."""
    L = _lit(prob, 5)
    Y = np.vstack([prob.Y, prob.replicate_data])
    alpha = L["fullalpha"][:5]
    Rm = np.random.default_rng(0).normal(size=(5, 5)) + 3 * np.eye(5)
    W2, newY2 = O._w_newy(Y, Y.shape[0], Rm @ alpha, prob.ctl)
    assert O.rel_frob(newY2, L["newY"]) < 1e-10
    ac = alpha[:, prob.ctl]
    A = Y[:, prob.ctl] @ ac.T
    A = A - A.mean(axis=0, keepdims=True)
    Lc = np.linalg.cholesky(ac @ ac.T)
    At = np.linalg.solve(Lc, A.T).T
    at = np.linalg.solve(Lc, alpha)
    for k in (1, 2, 5):
        ref = _lit(prob, k)["newY"]
        assert O.rel_frob(Y - At[:, :k] @ at[:k], ref) < 1e-11
        Wk = At[:, :k] @ np.linalg.inv(Lc)[:k, :k]
        assert O.rel_frob(Wk @ alpha[:k], _lit(prob, k)["W"] @ _lit(prob, k)["fullalpha"][:k]) < 1e-9


def test_user_fullalpha_and_k0_and_no_replicates(prob):
    L = _lit(prob, 3)
    again = _lit(prob, 2, fullalpha=L["fullalpha"])                       # ruvIII.R:138
    assert O.rel_frob(again["newY"], _lit(prob, 2)["newY"]) < 1e-12
    Y = np.vstack([prob.Y, prob.replicate_data])
    assert np.array_equal(_lit(prob, 0)["newY"], Y)                      # ruvIII.R:134-136
    uniq = ["u%d" % i for i in range(len(prob.replicate_names))]
    nr = O.ruvIII_literal(prob.assay, prob.sample_names, prob.replicate_data, uniq, prob.ctl, 3, apply_log=False)
    assert np.array_equal(nr["newY"], Y) and nr["W"] is None               # ruvIII.R:128-129
    with pytest.raises(ValueError):
        O.ruvIII_literal(prob.assay, prob.sample_names, prob.replicate_data, prob.replicate_names, prob.ctl, None)
    with pytest.raises(ValueError):                                        # fastruvIII.R:67-68
        O.fastruvIII_literal(prob.assay, prob.replicate_data, uniq, prob.ctl, 2, apply_log=False)


def test_dense_design_gene_side_route():
    """Rows > genes: the structural oracle takes the gene-side Gram (appendix A-5) and still matches the literal."""
    Y, gid, ctl = make_dense_design(m=90, n=40, group_size=3, n_ctl=12, k=3, seed=5)
    names = ["g%d" % g for g in gid]
    assay = Y.T
    L = O.ruvIII_literal(assay, names, np.zeros((0, 40)), [], ctl, 3, apply_log=False)
    S = O.ruvIII_structural(assay, names, np.zeros((0, 40)), [], ctl, 3, apply_log=False)
    assert O.rel_frob(S["newY"], L["newY"]) < 1e-10


def test_apply_log_and_average_rep(prob):
    counts = np.round(np.exp2(prob.assay) - 1).clip(min=0)
    a = O.ruvIII_literal(counts, prob.sample_names, prob.replicate_data, prob.replicate_names, prob.ctl, 2,
                         apply_log=True, pseudo_count=1.0)
    b = O.ruvIII_literal(np.log2(counts + 1.0), prob.sample_names, prob.replicate_data, prob.replicate_names,
                         prob.ctl, 2, apply_log=False)
    assert np.array_equal(a["newY"], b["newY"])
    avg = O.ruvIII_literal(prob.assay, prob.sample_names, prob.replicate_data, prob.replicate_names, prob.ctl, 2,
                           apply_log=False, apply_average_rep=True)
    assert avg["newY"].shape[0] == 150 + 11                              # one row per column of M (ruvIII.R:149)


# ---- the second, independent oracle (oracle/ruviii_oracle2.py): different numerical routes, no shared arithmetic ----------

def _o2():
    from oracle import ruviii_oracle2 as O2
    return O2


@pytest.mark.parametrize("shape", [dict(m1=60, n=300, groups=7, rep=3, n_ctl=40, k=4, seed=3),
                                   dict(m1=150, n=700, groups=11, rep=3, n_ctl=80, k=5, seed=7),
                                   dict(m1=90, n=400, groups=20, rep=2, n_ctl=60, k=9, seed=11)])
def test_independent_oracle_agrees_with_the_restatement(shape):
    """QR-projection residop, gesvd of Y0 itself, least-squares W and RUV1 (oracle2) against the R-statement transcription
    with dgesdd / explicit inverses (oracle): newY, W alpha (sign-free), fullalpha rows after sign alignment."""
    O2 = _o2()
    p = make_problem(**shape)
    k = shape["k"]
    a = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False)
    b = O2.ruvIII(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False)
    assert O.rel_frob(b["newY"], a["newY"]) < 1e-10
    assert b["fullalpha"].shape == a["fullalpha"].shape
    assert O.rel_frob(O.align_rows_sign(b["fullalpha"][:k], a["fullalpha"][:k]), a["fullalpha"][:k]) < 1e-8
    assert O.rel_frob(b["W"] @ b["fullalpha"][:k], a["W"] @ a["fullalpha"][:k]) < 1e-10
    fa = O.fastruvIII_literal(p.assay, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False)
    fb = O2.fastruvIII(p.assay, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False)
    assert O.rel_frob(fb["newY"], fa["newY"]) < 1e-10
    assert O.rel_frob(O.align_rows_sign(fb["fullalpha"], fa["fullalpha"]), fa["fullalpha"]) < 1e-8


def test_independent_oracle_options(golden_mid):
    """apply.log, apply.average.rep, eta (RUV1 with the intercept rule of design.matrix), index-vector ctl."""
    O2 = _o2()
    g = golden_mid
    m = g["meta"]
    k = m["k"]
    args = (m["sample_names"], g["replicate_data"], m["replicate_names"])
    b = O2.ruvIII(g["counts"], *args, g["ctl"], k, apply_log=True, pseudo_count=1.0)
    assert O.rel_frob(b["newY"], g["ruvIII.log.newY"]) < 1e-10
    b = O2.ruvIII(g["assay"], *args, np.asarray(m["ctl_index_0based"]), k, apply_log=False)          # tological()
    assert O.rel_frob(b["newY"], g["ruvIII.newY"]) < 1e-10
    eta = g["eta"]                                                                                   # n x 1, as R passes it
    b = O2.ruvIII(g["assay"], *args, g["ctl"], k, apply_log=False, eta=eta, include_intercept=True)
    assert O.rel_frob(b["newY"], g["ruvIII.eta.newY"]) < 1e-10
    a = O.ruvIII_literal(g["assay"], *args, g["ctl"], k, apply_log=False, apply_average_rep=True)
    b = O2.ruvIII(g["assay"], *args, g["ctl"], k, apply_log=False, apply_average_rep=True)
    assert O.rel_frob(b["newY"], a["newY"]) < 1e-10
    # an eta that already spans the constant vector gets no second intercept (design.matrix), a scalar means intercept only
    n = g["assay"].shape[0]
    eta2 = np.vstack([np.full(n, 3.0), np.sin(np.arange(n))])
    a = O.ruvIII_literal(g["assay"], *args, g["ctl"], k, apply_log=False, eta=eta2)
    b = O2.ruvIII(g["assay"], *args, g["ctl"], k, apply_log=False, eta=eta2)
    assert O.rel_frob(b["newY"], a["newY"]) < 1e-10
    assert O.design_matrix_t(eta2, n).shape[0] == 2 and O.design_matrix_t(1.0, n).shape == (1, n)
    assert O.design_matrix_t(np.sin(np.arange(n)), n).shape[0] == 2


def test_every_golden_fixture_reproduces_under_both_oracles(golden_tiny, golden_mid):
    O2 = _o2()
    for g in (golden_tiny, golden_mid):
        m = g["meta"]
        args = (g["assay"], m["sample_names"], g["replicate_data"], m["replicate_names"], g["ctl"])
        for fn in (lambda k: O.ruvIII_literal(*args, k, apply_log=False), lambda k: O2.ruvIII(*args, k, apply_log=False)):
            assert O.rel_frob(fn(m["k"])["newY"], g["ruvIII.newY"]) < 1e-10
            for k in m["ks_multi"]:
                assert O.rel_frob(fn(k)["newY"], g["ruvIIIMultipleK.k%d.newY" % k]) < 1e-10
        f2 = O2.fastruvIII(g["assay"], g["replicate_data"], m["replicate_names"], g["ctl"], m["k"], apply_log=False)
        assert O.rel_frob(f2["newY"], g["fastruvIII.newY"]) < 1e-10


def test_replay_script_covers_every_fixture():
    """tests/golden/replay.R is the one-command R pin: it must read every stored output named in the manifests."""
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    script = open(os.path.join(here, "replay.R")).read()
    for case in ("tiny", "mid"):
        names = [ln.split("\t")[0] for ln in open(os.path.join(here, case + ".manifest.tsv")).read().splitlines()[1:]]
        for name in names:
            if name in ("k", "k_multi"):
                continue
            key = "ruvIIIMultipleK.k" if name.startswith("ruvIIIMultipleK.k") else name
            assert ('"%s' % key) in script, "replay.R never reads %s.%s.bin" % (case, name)
        for side in ("sample_names", "replicate_names", "ctl_index_1based"):
            assert os.path.exists(os.path.join(here, "%s.%s.txt" % (case, side)))


def test_structural_rows_matches_full_structural(prob):
    """The row-subset form bench.py uses for its parity figure equals the full structural restatement."""
    p = prob
    rows = np.array([0, 3, 17, 88, 149])
    sub, lam, fa = O.ruvIII_structural_rows(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, [0, 1, 3, 5], rows,
                                            apply_log=False)
    full = O.ruvIIIMultipleK_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, [0, 1, 3, 5],
                                        apply_log=False)
    for k in (0, 1, 3, 5):
        assert O.rel_frob(sub[k], full[k]["newY"][rows]) < 1e-12


def test_side_fixtures_reproduce_from_the_oracles(golden_mid):
    """The base-R-checkable fixtures of the steps either side of the path (replay.R, first part): regenerated here from the
    numpy restatements, and cross-checked by routes that share no code with them (scipy's F test, numpy's corrcoef)."""
    from scipy import stats
    from oracle import ncg_oracle as NO
    from oracle import pca_oracle as PCAO
    g = golden_mid
    assay, batch, cov = g["assay"], np.asarray(g["meta"]["side_batch"]), g["side.covariate"][:, 0]
    cd = {"batch": list(batch), "cov": list(cov)}
    assert O.rel_frob(NO.lm_residuals(assay, cd, ["batch", "cov"]), g["side.lm_residuals"]) < 1e-12
    F = np.array([stats.f_oneway(*[row[batch == b] for b in np.unique(batch)]).statistic for row in assay])
    assert O.rel_frob(F, g["side.anovaF"][:, 0]) < 1e-10
    r = np.array([np.corrcoef(row, cov)[0, 1] for row in assay])
    assert np.max(np.abs(r - g["side.pearson"][:, 0])) < 1e-12
    rho = np.array([stats.spearmanr(row, cov).statistic for row in assay])
    assert np.max(np.abs(rho - g["side.spearman"][:, 0])) < 1e-12
    sv = PCAO.fast_pca(assay, 60, apply_log=False)
    assert np.allclose(sv["d"][:59], g["side.svd_d"][:59, 0], rtol=1e-12)
    # residuals are orthogonal to the model matrix (intercept, batch dummies, covariate)
    D = np.stack([np.ones(60)] + [(batch == b).astype(float) for b in np.unique(batch)[1:]] + [cov], axis=1)
    assert np.max(np.abs(g["side.lm_residuals"] @ D)) < 1e-9
