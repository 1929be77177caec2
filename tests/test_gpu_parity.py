"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): relative Frobenius error <= 1e-9 on newY; fullalpha rows / W columns compared
after sign alignment; W %*% alpha compared sign-free.  PARITY UNPINNED against R itself (see the oracle header).
"""
import dataclasses
import os

import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api
from ruviii_b200.synth import make_config, make_dense_design, make_problem

pytestmark = pytest.mark.gpu

TOL_NEWY = 1e-9          # north-star tolerance on newY
TOL_SIGNED = 1e-7        # sign-aligned alpha rows / W columns (conditioned by the eigen-gaps of the synthetic data)


def run_engine(eng, p, ks, mode=_cabi.MODE_STACKED, apply_log=False, pseudo_count=1.0, want_all=True, n_alpha_rows=0,
               assay=None):
    assay = p.assay if assay is None else assay
    names = list(p.replicate_names) if mode == _cabi.MODE_FAST else list(p.sample_names) + list(p.replicate_names)
    groups = api.group_codes(names)
    prm = _cabi.make_params(mode=mode, apply_log=apply_log, pseudo_count=pseudo_count, n_alpha_rows=n_alpha_rows,
                            want_ps_rows=(mode == _cabi.MODE_STACKED))
    return eng.normalise(prm, assay.T, p.replicate_data, groups, p.ctl, ks, want_fullalpha=want_all, want_W=want_all,
                         want_ps=(mode == _cabi.MODE_STACKED))


def check_against(res, i, ref, k, m1, stacked=True):
    newY = np.vstack([res["newY"][i], res["newY_ps"][i]]) if stacked else res["newY"][i]
    assert O.rel_frob(newY, ref["newY"]) < TOL_NEWY
    if k == 0 or ref["W"] is None:
        return
    kk = ref["W"].shape[1]
    fa = O.align_rows_sign(res["fullalpha"][:kk], ref["fullalpha"][:kk])
    assert O.rel_frob(fa, ref["fullalpha"][:kk]) < TOL_SIGNED
    W = O.align_rows_sign(res["W"][i].T, ref["W"].T).T
    assert O.rel_frob(W, ref["W"]) < TOL_SIGNED
    assert O.rel_frob(res["W"][i] @ res["fullalpha"][:kk], ref["W"] @ ref["fullalpha"][:kk]) < TOL_NEWY * 10


# ---- golden fixtures through the R-shaped API ---------------------------------------------------------------

@pytest.mark.parametrize("case", ["tiny", "mid"])
def test_golden_through_r_api(engine, case, golden_tiny, golden_mid):
    g = golden_tiny if case == "tiny" else golden_mid
    m = g["meta"]
    api.set_engine(engine)
    se = rv.SummarizedExperiment(assays={"HTseq": np.asfortranarray(g["assay"])}, col_names=m["sample_names"])
    rd = rv.ReplicateData(g["replicate_data"], m["replicate_names"])
    k = m["k"]
    out = rv.ruvIII(se, "HTseq", apply_log=False, replicate_data=rd, ctl=g["ctl"], k=k, return_info=True,
                    save_se_obj=False, verbose=False)
    assert O.rel_frob(out["newY"], g["ruvIII.newY"]) < TOL_NEWY
    assert out["fullalpha"].shape == g["ruvIII.fullalpha"].shape            # r rows, ruvIII.R:140
    assert O.rel_frob(O.align_rows_sign(out["fullalpha"][:k], g["ruvIII.fullalpha"][:k]), g["ruvIII.fullalpha"][:k]) < TOL_SIGNED
    assert O.rel_frob(out["W"] @ out["fullalpha"][:k], g["ruvIII.W"] @ g["ruvIII.fullalpha"][:k]) < 1e-8
    # save.se.obj = TRUE: new assay RUV_K<k>_on_<assay> = t(newY[1:m1, ])   (ruvIII.R:154-158)
    se2 = rv.ruvIII(se, "HTseq", apply_log=False, replicate_data=rd, ctl=g["ctl"], k=k, verbose=False)
    name = "RUV_K%d_on_HTseq" % k
    m1 = len(m["sample_names"])
    assert se2.assays[name].shape == g["assay"].shape
    assert O.rel_frob(se2.assays[name], g["ruvIII.newY"][:m1].T) < TOL_NEWY
    # t(newY) return (ruvIII.R:191-192)
    t_newY = rv.ruvIII(se, "HTseq", apply_log=False, replicate_data=rd, ctl=g["ctl"], k=k, save_se_obj=False, verbose=False)
    assert O.rel_frob(t_newY, g["ruvIII.newY"].T) < TOL_NEWY
    # fastruvIII
    f = rv.fastruvIII(se, "HTseq", apply_log=False, replicate_data=rd, ctl=g["ctl"], k=k, return_info=True,
                      save_se_obj=False, verbose=False)
    assert O.rel_frob(f["newY"], g["fastruvIII.newY"]) < TOL_NEWY
    assert f["fullalpha"].shape == g["fastruvIII.fullalpha"].shape          # k rows, fastruvIII.R:171
    # ruvIIIMultipleK, SE branch (ruvIIIMultipleK.R:75-94)
    se3 = rv.ruvIIIMultipleK(se, "HTseq", apply_log=False, replicate_data=rd, ctl=g["ctl"], k=m["ks_multi"], verbose=False)
    for kv in m["ks_multi"]:
        assert O.rel_frob(se3.assays["RUV_K%d_on_HTseq" % kv], g["ruvIIIMultipleK.k%d.newY" % kv][:m1].T) < TOL_NEWY
    lst = rv.ruvIIIMultipleK(se, "HTseq", apply_log=False, replicate_data=rd, ctl=g["ctl"], k=m["ks_multi"],
                             save_se_obj=False, verbose=False)
    assert list(lst)[0].startswith("RUVIII_K:") and len(lst) == len(m["ks_multi"])
    if case == "mid":                                                        # apply.log = TRUE on counts
        sec = rv.SummarizedExperiment(assays={"HTseq": np.asfortranarray(g["counts"])}, col_names=m["sample_names"])
        o = rv.ruvIII(sec, "HTseq", apply_log=True, pseudo_count=1, replicate_data=rd, ctl=g["ctl"], k=k,
                      return_info=True, save_se_obj=False, verbose=False)
        assert O.rel_frob(o["newY"], g["ruvIII.log.newY"]) < TOL_NEWY


# ---- seeded shapes against the literal oracle -----------------------------------------------------------------

SHAPES = [
    # m1, n, groups, rep, n_ctl, ks
    (40, 64, 5, 2, 16, [1]),
    (90, 301, 8, 3, 40, [3]),              # odd gene count (pitch padding, unaligned host rows)
    (200, 1001, 45, 3, 120, [1, 4, 9]),    # R = 135 > one 128-row Gram tile; multi-k
    (1500, 2050, 30, 9, 200, [2, 17]),     # groups of 9 (two-pass residop path), k > 16 (32-wide register tile)
    (64, 4100, 129, 2, 300, [5, 5, 7]),    # R = 258: three tile rows; duplicate k
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "m%d_n%d_g%dx%d" % s[:4])
def test_stacked_matches_literal(engine, shape):
    m1, n, groups, rep, n_ctl, ks = shape
    p = make_problem(m1=m1, n=n, groups=groups, rep=rep, n_ctl=n_ctl, k=max(ks), seed=1000 + n)
    res = run_engine(engine, p, ks)
    info = res["info"]
    assert info.p == m1 + groups and info.active_rows == groups * rep and info.r == min(groups * (rep - 1), int(p.ctl.sum()))
    for i, k in enumerate(ks):
        ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False)
        check_against(res, i, ref, k, m1)


@pytest.mark.parametrize("shape", SHAPES[:4], ids=lambda s: "m%d_n%d_g%dx%d" % s[:4])
def test_fast_matches_literal(engine, shape):
    m1, n, groups, rep, n_ctl, ks = shape
    p = make_problem(m1=m1, n=n, groups=groups, rep=rep, n_ctl=n_ctl, k=max(ks), seed=2000 + n)
    for k in ks[:2]:
        res = run_engine(engine, p, [k], mode=_cabi.MODE_FAST)
        ref = O.fastruvIII_literal(p.assay, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False)
        check_against(res, 0, ref, k, m1, stacked=False)
        assert res["fullalpha"].shape[0] == k and res["W"][0].shape == (m1, k)


def test_full_fullalpha_rows(engine):
    """return.info = TRUE keeps all r = min(m - ncol(M), sum(ctl)) rows of fullalpha (ruvIII.R:140,187-188)."""
    # enough samples per (biology, batch) cell that no two pseudo-samples coincide: rank(Y0) = r exactly.  (With
    # a rank-deficient Y0 the reference's trailing rows t(U) %*% Y are arbitrary null-space vectors times Y.)
    p = make_problem(m1=400, n=500, groups=12, rep=3, n_ctl=60, k=3, seed=31)
    res = run_engine(engine, p, [3], n_alpha_rows=-1)
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 3, apply_log=False)
    assert res["fullalpha"].shape == ref["fullalpha"].shape == (24, 500)
    # rows with well separated singular values agree after sign alignment; all rows span the same space
    top = 10
    fa = O.align_rows_sign(res["fullalpha"][:top], ref["fullalpha"][:top])
    assert O.rel_frob(fa, ref["fullalpha"][:top]) < 1e-6
    G1 = res["fullalpha"] @ res["fullalpha"].T
    G2 = ref["fullalpha"] @ ref["fullalpha"].T
    assert abs(np.trace(G1) - np.trace(G2)) < 1e-9 * np.trace(G2)


def test_apply_log_on_counts(engine):
    p = make_problem(m1=70, n=333, groups=6, rep=3, n_ctl=50, k=2, seed=44, counts=True)
    res = run_engine(engine, p, [2], apply_log=True, pseudo_count=1.0)
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 2, apply_log=True,
                           pseudo_count=1.0)
    check_against(res, 0, ref, 2, 70)


# ---- edge cases the reference handles -------------------------------------------------------------------------

def test_k0_and_no_replicates_return_Y(engine):
    p = make_problem(m1=33, n=130, groups=4, rep=2, n_ctl=20, k=2, seed=5)
    res = run_engine(engine, p, [0, 2], want_all=False)
    assert np.array_equal(res["newY"][0], p.Y)                                    # ruvIII.R:134-136
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 2, apply_log=False)
    assert O.rel_frob(res["newY"][1], ref["newY"][:33]) < TOL_NEWY
    uniq = ["u%d" % i for i in range(len(p.replicate_names))]
    groups = api.group_codes(p.sample_names + uniq)
    out = engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl, [2])
    assert out["info"].no_replicates == 1 and np.array_equal(out["newY"][0], p.Y)  # ruvIII.R:128-129


def test_k_larger_than_r_is_clamped(engine):
    """alpha <- fullalpha[1:min(k, nrow(fullalpha)), ]  (ruvIII.R:142)."""
    p = make_problem(m1=50, n=220, groups=3, rep=2, n_ctl=30, k=2, seed=6)         # r = 3
    res = run_engine(engine, p, [3, 8])
    assert res["info"].r == 3 and res["info"].kmax_eff == 3
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 8, apply_log=False)
    check_against(res, 1, ref, 8, 50)
    assert np.array_equal(res["newY"][0], res["newY"][1])


def test_replicated_samples_general_M(engine):
    """Duplicate sample ids become replicate groups: general M, not only PRPS rows."""
    p = make_problem(m1=60, n=400, groups=5, rep=3, n_ctl=40, k=3, seed=8)
    names = list(p.sample_names)
    for i in range(0, 20, 2):
        names[i + 1] = names[i]                                                    # 10 technical-replicate pairs
    names[30] = p.replicate_names[0]                                               # a sample joins a PRPS set
    groups = api.group_codes(names + list(p.replicate_names))
    prm = _cabi.make_params(want_ps_rows=True)
    res = engine.normalise(prm, p.Y, p.replicate_data, groups, p.ctl, [3], want_fullalpha=True, want_W=True, want_ps=True)
    ref = O.ruvIII_literal(p.assay, names, p.replicate_data, p.replicate_names, p.ctl, 3, apply_log=False)
    assert res["info"].active_rows == 20 + 15 + 1
    check_against(res, 0, ref, 3, 60)


def test_dense_design_sample_side(engine):
    """Every row in a replicate group of 3, no PRPS rows (SURVEY.md 8d, C3d at small scale)."""
    Y, gid, ctl = make_dense_design(m=300, n=900, group_size=3, n_ctl=100, k=4, seed=9)
    res = engine.normalise(_cabi.make_params(), Y, np.zeros((0, 900)), gid, ctl, [4], want_fullalpha=True, want_W=True)
    names = ["g%d" % g for g in gid]
    ref = O.ruvIII_literal(Y.T, names, np.zeros((0, 900)), [], ctl, 4, apply_log=False)
    assert res["info"].active_rows == 300
    check_against(res, 0, ref, 4, 300, stacked=False)


def test_gene_side_gram(engine):
    """More non-singleton rows than genes (BASELINE config 5's regime at small scale): the engine decomposes
    t(Z) %*% Z (genes x genes) and takes alpha = diag(sigma) V' (SURVEY.md appendix A-5)."""
    Y, gid, ctl = make_dense_design(m=600, n=150, group_size=3, n_ctl=60, k=4, seed=17)
    res = engine.normalise(_cabi.make_params(), Y, np.zeros((0, 150)), gid, ctl, [2, 4], want_fullalpha=True, want_W=True)
    assert res["info"].gram_side == 2 and res["info"].gram_order == 150 and res["info"].active_rows == 600
    names = ["g%d" % g for g in gid]
    for i, k in enumerate([2, 4]):
        ref = O.ruvIII_literal(Y.T, names, np.zeros((0, 150)), [], ctl, k, apply_log=False)
        check_against(res, i, ref, k, 600, stacked=False)
    # forcing the gene side on an ordinary PRPS problem gives the sample-side answer
    p = make_problem(m1=120, n=900, groups=40, rep=3, n_ctl=80, k=5, seed=18)
    groups = api.group_codes(p.sample_names + p.replicate_names)
    a = engine.normalise(_cabi.make_params(gram_side=1), p.Y, p.replicate_data, groups, p.ctl, [5])
    b = engine.normalise(_cabi.make_params(gram_side=2), p.Y, p.replicate_data, groups, p.ctl, [5])
    assert a["info"].gram_side == 1 and b["info"].gram_side == 2 and b["info"].gram_order == 900
    assert O.rel_frob(b["newY"][0], a["newY"][0]) < 1e-10


def test_apply_average_rep(engine):
    """apply.average.rep = TRUE (ruvIII.R:148-149): one row per column of M, the mean of the group's rows of newY."""
    p = make_problem(m1=50, n=260, groups=6, rep=3, n_ctl=35, k=3, seed=19)
    names = list(p.sample_names)
    names[7] = names[3]                                              # one pair of technical replicates among the samples
    api.set_engine(engine)
    se = rv.SummarizedExperiment(assays={"a": p.assay}, col_names=names)
    rd = rv.ReplicateData(p.replicate_data, p.replicate_names)
    out = rv.ruvIII(se, "a", apply_log=False, replicate_data=rd, ctl=p.ctl, k=3, apply_average_rep=True,
                    return_info=True, save_se_obj=False, verbose=False)
    ref = O.ruvIII_literal(p.assay, names, p.replicate_data, p.replicate_names, p.ctl, 3, apply_log=False,
                           apply_average_rep=True)
    assert out["newY"].shape == ref["newY"].shape == (49 + 6, 260)
    assert O.rel_frob(out["newY"], ref["newY"]) < TOL_NEWY
    with pytest.raises(rv.RuviiiError):                            # fastruvIII: t(M) (n_ps rows) x newY (m1 rows) does not conform
        rv.fastruvIII(se, "a", apply_log=False, replicate_data=rd, ctl=p.ctl, k=3, apply_average_rep=True, verbose=False)


def test_user_supplied_fullalpha(engine):
    """fullalpha argument skips residop / Gram / SVD (ruvIII.R:138)."""
    p = make_problem(m1=45, n=260, groups=6, rep=3, n_ctl=35, k=4, seed=10)
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 4, apply_log=False)
    api.set_engine(engine)
    se = rv.SummarizedExperiment(assays={"a": p.assay}, col_names=p.sample_names)
    rd = rv.ReplicateData(p.replicate_data, p.replicate_names)
    out = rv.ruvIII(se, "a", apply_log=False, replicate_data=rd, ctl=p.ctl, k=2, fullalpha=ref["fullalpha"],
                    return_info=True, save_se_obj=False, verbose=False)
    ref2 = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 2, apply_log=False,
                            fullalpha=ref["fullalpha"])
    assert O.rel_frob(out["newY"], ref2["newY"]) < TOL_NEWY
    assert O.rel_frob(out["W"], ref2["W"]) < 1e-8


@pytest.mark.parametrize("mode", ["ruvIII", "fastruvIII"])
def test_eta_ruv1_prestep(engine, mode):
    """eta != NULL (ruvIII.R:116, fastruvIII.R:133-134): ruv::RUV1 removes gene-wise covariates through the control genes
    from every row of Y and replicate.data before anything else."""
    p = make_problem(m1=150, n=1203, groups=20, rep=3, n_ctl=150, k=4, seed=611)
    rng = np.random.default_rng(5)
    eta = np.vstack([rng.normal(size=1203), np.linspace(-1, 1, 1203) ** 2])          # two gene-wise covariates (GC, length ...)
    se = rv.SummarizedExperiment(assays={"logcounts": p.assay}, col_names=p.sample_names)
    rd = rv.ReplicateData(p.replicate_data, p.replicate_names)
    if mode == "ruvIII":
        out = rv.ruvIII(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=4, eta=eta.T, return_info=True,
                        save_se_obj=False, verbose=False)
        ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 4, apply_log=False, eta=eta)
        plain = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 4, apply_log=False)
    else:
        out = rv.fastruvIII(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=4, eta=eta.T, return_info=True,
                            save_se_obj=False, verbose=False)
        ref = O.fastruvIII_literal(p.assay, p.replicate_data, p.replicate_names, p.ctl, 4, apply_log=False, eta=eta)
        plain = O.fastruvIII_literal(p.assay, p.replicate_data, p.replicate_names, p.ctl, 4, apply_log=False)
    assert O.rel_frob(out["newY"], ref["newY"]) < TOL_NEWY
    assert O.rel_frob(plain["newY"], ref["newY"]) > 1e-4                              # eta really changes the answer
    assert O.rel_frob(out["W"] @ out["fullalpha"][:4], ref["W"] @ ref["fullalpha"][:4]) < TOL_NEWY * 10
    # eta is cleared again: the next call without eta is the plain result
    again = (rv.ruvIII if mode == "ruvIII" else rv.fastruvIII)(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl,
                                                             k=4, return_info=True, save_se_obj=False, verbose=False)
    assert O.rel_frob(again["newY"], plain["newY"]) < TOL_NEWY
    # k = 0 still applies RUV1 (ruvIII.R:116 runs before :134-136); intercept only = eta given as a single number
    k0 = rv.ruvIII(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=0, eta=1, save_se_obj=False, verbose=False)
    ref0 = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 0, apply_log=False,
                            eta=np.zeros((0, 1203)))
    assert O.rel_frob(k0.T, ref0["newY"]) < TOL_NEWY
    # collinear covariates: solve() in RUV1 fails
    with pytest.raises(rv.RuviiiError) as e:
        rv.ruvIII(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=2, eta=np.vstack([eta[0], eta[0]]).T,
                  save_se_obj=False, verbose=False)
    assert e.value.code == _cabi.ERR_SINGULAR


def test_singular_control_block_is_an_error(engine):
    """solve(ac %*% t(ac)) fails in R when the control genes carry no unwanted variation (ruvIII.R:144)."""
    p = make_problem(m1=40, n=120, groups=4, rep=2, n_ctl=15, k=2, seed=11)
    PS = p.replicate_data.copy()
    PS[:, p.ctl] = 7.0                                             # residuals vanish on ctl => ac == 0
    groups = api.group_codes(p.sample_names + p.replicate_names)
    with pytest.raises(rv.RuviiiError) as e:
        engine.normalise(_cabi.make_params(), p.Y, PS, groups, p.ctl, [2])
    assert e.value.code == _cabi.ERR_SINGULAR


def test_dimension_errors(engine):
    p = make_problem(m1=30, n=60, groups=3, rep=2, n_ctl=10, k=1, seed=12)
    groups = api.group_codes(p.sample_names + p.replicate_names)
    with pytest.raises(rv.RuviiiError):
        engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups[:-1], p.ctl, [1])
    with pytest.raises(rv.RuviiiError):
        engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl[:-1], [1])
    with pytest.raises(rv.RuviiiError):
        engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl, [-1])
    # k = 40 is clamped to r = 3 (ruvIII.R:142), so it is legal here; an EFFECTIVE k > 32 is what is unsupported
    ok = engine.normalise(_cabi.make_params(), p.Y, p.replicate_data, groups, p.ctl, [40])
    assert ok["info"].kmax_eff == 3
    big = make_problem(m1=400, n=700, groups=40, rep=3, n_ctl=100, k=40, seed=15)            # r = 80
    with pytest.raises(rv.RuviiiError) as e:
        engine.normalise(_cabi.make_params(), big.Y, big.replicate_data,
                         api.group_codes(big.sample_names + big.replicate_names), big.ctl, [40])
    assert e.value.code == _cabi.ERR_UNSUPPORTED


def test_fast_residop_primitive(engine):
    """RcppExports.R:4-6 fastResidop(A, B) with B a replicate matrix == fast.residop (fastruvIII.R:88-95)."""
    rng = np.random.default_rng(3)
    A = rng.normal(size=(37, 211))
    gid = rng.integers(0, 9, size=37).astype(np.int32)
    out = engine.fast_residop(A, gid)
    M = O.replicate_matrix(gid)
    assert O.rel_frob(out, O.residop(A, M)) < 1e-13 or np.abs(out - O.residop(A, M)).max() < 1e-12


def test_gram_kernels_agree():
    """The DMMA Gram (impl 1) and the plain DFMA cross-check kernel (impl 0) give the same newY and eigenvalues."""
    p = make_problem(m1=100, n=3000, groups=50, rep=3, n_ctl=150, k=6, seed=13)
    outs = {}
    for impl in (0, 1):
        os.environ["RUVIII_GRAM_IMPL"] = str(impl)
        try:
            with rv.Engine(devices=[0]) as eng:
                res = run_engine(eng, p, [6])
                outs[impl] = (res["newY"][0].copy(), eng.eigenvalues(7))
        finally:
            os.environ.pop("RUVIII_GRAM_IMPL", None)
    assert O.rel_frob(outs[1][0], outs[0][0]) < 1e-11
    assert np.allclose(outs[1][1], outs[0][1], rtol=1e-12)
    S = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 6, apply_log=False)
    assert np.allclose(outs[1][1][:6], S["eigenvalues"][:6], rtol=1e-11)


@pytest.mark.parametrize("env", [{"RUVIII_GRAM_IMPL": "2"}, {"RUVIII_GRAM_IMPL": "4"}, {"RUVIII_UPD_IMPL": "2"},
                                 {"RUVIII_UPD_IMPL": "1"}, {"RUVIII_UPD_PREFETCH": "0"}],
                         ids=["gram_tma", "gram_128", "update_tma", "update_regs", "update_noprefetch"])
def test_tma_kernel_variants_match(env):
    """The TMA + mbarrier variants of the Gram and of the update give the same results as the default kernels:
    odd gene count, R = 258 (three tile rows, last one ragged), multi-k, k = 0 included."""
    p = make_problem(m1=333, n=4101, groups=129, rep=2, n_ctl=300, k=7, seed=16)
    ks = [0, 2, 7]
    with rv.Engine(devices=[0]) as eng:
        base = run_engine(eng, p, ks)
    os.environ.update(env)
    try:
        with rv.Engine(devices=[0]) as eng:
            alt = run_engine(eng, p, ks)
    finally:
        for k in env:
            os.environ.pop(k, None)
    for i in range(len(ks)):
        assert O.rel_frob(alt["newY"][i], base["newY"][i]) < 1e-12
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 7, apply_log=False)
    check_against(alt, 2, ref, 7, 333)
    assert np.array_equal(alt["newY"][0], p.Y)


@pytest.mark.parametrize("apply_log", [False, True], ids=["log2_input", "counts_apply_log"])
@pytest.mark.parametrize("zc_pct", ["", "0", "40", "100"])
def test_pipelined_host_path_matches_sequential(engine, apply_log, zc_pct, monkeypatch):
    """RUVIII_PIPELINE=3: the round-1 pipeline on page-locked buffers (control columns gathered on the device for the
    leading row chunks and by zero-copy for the trailing ones, chunked update); whatever the split, it must return exactly
    what the sequential upload -> run -> download path returns."""
    if zc_pct:
        monkeypatch.setenv("RUVIII_ZC_PCT", zc_pct)
    p = make_problem(m1=700, n=3001, groups=60, rep=3, n_ctl=200, k=7, seed=20, counts=apply_log)
    ks = [0, 3, 7]
    groups = api.group_codes(p.sample_names + p.replicate_names)
    prm = _cabi.make_params(apply_log=apply_log, pseudo_count=1.0)
    seq = engine.normalise(prm, p.Y, p.replicate_data, groups, p.ctl, ks)
    assert seq["info"].pipelined == 0
    Yp, outp = _cabi.PinnedArray((700, 3001)), _cabi.PinnedArray((3, 700, 3001))
    Yp.array[...] = p.Y
    outp.array[...] = -1.0
    monkeypatch.setenv("RUVIII_PIPELINE", "3")
    info = engine.normalise_oneshot(prm, Yp.array, p.replicate_data, groups, p.ctl, ks, outp.array)
    assert info.pipelined == 1
    for i in range(3):
        assert np.array_equal(outp.array[i], seq["newY"][i])
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 7, apply_log=apply_log,
                           pseudo_count=1.0)
    assert O.rel_frob(outp.array[2], ref["newY"][:700]) < TOL_NEWY


@pytest.mark.parametrize("apply_log", [False, True], ids=["log2_input", "counts_apply_log"])
@pytest.mark.parametrize("threads", ["1", "5"])
@pytest.mark.parametrize("m1", [700, 2500])
def test_staged_host_path_matches_sequential(engine, apply_log, threads, m1, monkeypatch):
    """The default one-shot path (ruviii_normalise / ruviii_normalise_blocks): control columns copied out by host threads,
    assay streamed through in row chunks -- from PAGEABLE memory through the pinned bounce buffers (what R hands over),
    from pinned memory directly, and with one separately allocated output per k.  Bit-identical to the sequential path."""
    monkeypatch.setenv("RUVIII_HOST_THREADS", threads)
    n = 3001 if m1 == 700 else 20000                          # 2,500 x 20,000: several row chunks per slot
    p = make_problem(m1=m1, n=n, groups=60, rep=3, n_ctl=200, k=7, seed=20, counts=apply_log)
    ks = [0, 3, 7]
    groups = api.group_codes(p.sample_names + p.replicate_names)
    prm = _cabi.make_params(apply_log=apply_log, pseudo_count=1.0)
    monkeypatch.setenv("RUVIII_PIPELINE", "2")
    seq = engine.normalise(prm, p.Y, p.replicate_data, groups, p.ctl, ks)
    seq_one = np.full((3, m1, n), -1.0)
    info = engine.normalise_oneshot(prm, np.ascontiguousarray(p.Y), p.replicate_data, groups, p.ctl, ks, seq_one)
    assert info.pipelined == 0 and np.array_equal(seq_one, seq["newY"])
    monkeypatch.delenv("RUVIII_PIPELINE")
    # pageable in, pageable out (numpy heap)
    out = np.full((3, m1, n), -1.0)
    info = engine.normalise_oneshot(prm, np.ascontiguousarray(p.Y), p.replicate_data, groups, p.ctl, ks, out)
    assert info.pipelined == 3
    assert np.array_equal(out, seq["newY"])
    # one separately allocated matrix per k (the R shim's REALSXPs)
    outs = [np.full((m1, n), -2.0) for _ in ks]
    info = engine.normalise_blocks(prm, np.ascontiguousarray(p.Y), p.replicate_data, groups, p.ctl, ks, outs)
    assert info.pipelined == 3
    for i in range(3):
        assert np.array_equal(outs[i], seq["newY"][i])
    # pinned in, pinned out: no bounce buffers
    Yp, outp = _cabi.PinnedArray((m1, n)), _cabi.PinnedArray((3, m1, n))
    Yp.array[...] = p.Y
    outp.array[...] = -1.0
    info = engine.normalise_oneshot(prm, Yp.array, p.replicate_data, groups, p.ctl, ks, outp.array)
    assert info.pipelined == 2
    assert np.array_equal(outp.array, seq["newY"])
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 7, apply_log=apply_log,
                           pseudo_count=1.0) if m1 == 700 else O.ruvIII_structural(
        p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 7, apply_log=apply_log, pseudo_count=1.0)
    assert O.rel_frob(out[2], ref["newY"][:m1]) < TOL_NEWY


def test_eigensolvers_agree():
    """Block subspace iteration (solver 3) vs cuSOLVER syevdx (solver 2) vs the oracle, on a problem large enough
    for the iterative solver (R >= 128, rank >= 64).  newY depends only on the top-k subspace."""
    p = make_problem(m1=500, n=6000, groups=80, rep=3, n_ctl=300, k=8, seed=14)
    ref = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 8, apply_log=False)
    outs = {}
    for impl in (2, 3):
        os.environ["RUVIII_EIG_IMPL"] = str(impl)
        try:
            with rv.Engine(devices=[0]) as eng:
                res = run_engine(eng, p, [3, 8])
                assert res["info"].eig_solver == impl
                if impl == 3:
                    assert 0 < res["info"].eig_iterations <= 150
                outs[impl] = (np.vstack([res["newY"][1], res["newY_ps"][1]]), eng.eigenvalues(8), res["fullalpha"], res["W"][1])
        finally:
            os.environ.pop("RUVIII_EIG_IMPL", None)
    for impl in (2, 3):
        assert O.rel_frob(outs[impl][0], ref["newY"]) < TOL_NEWY
        assert np.allclose(outs[impl][1], ref["eigenvalues"][:8], rtol=1e-11)
        fa = O.align_rows_sign(outs[impl][2][:8], ref["fullalpha"][:8])
        assert O.rel_frob(fa, ref["fullalpha"][:8]) < TOL_SIGNED
    assert O.rel_frob(outs[3][0], outs[2][0]) < 1e-10
    # deterministic sign convention: both solvers return the same orientation
    assert O.rel_frob(outs[3][2], outs[2][2]) < TOL_SIGNED
    with rv.Engine(devices=[0]) as eng:                       # auto picks the iterative solver here
        assert run_engine(eng, p, [8])["info"].eig_solver == 3


# ---- BASELINE.json sizes ----------------------------------------------------------------------------------------

def test_config_C1_C2_full_size(engine):
    """C1/C2: 1,200 x 20,000, 60 PRPS sets x 3, 1,000 control genes, k = 5 -- both entry points, full oracle."""
    p = make_config("C1")
    res = run_engine(engine, p, [5], want_all=False)
    ref = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 5, apply_log=False)
    assert O.rel_frob(np.vstack([res["newY"][0], res["newY_ps"][0]]), ref["newY"]) < TOL_NEWY
    resf = run_engine(engine, p, [5], mode=_cabi.MODE_FAST, want_all=False)
    reff = O.fastruvIII_structural(p.assay, p.replicate_data, p.replicate_names, p.ctl, 5, apply_log=False)
    assert O.rel_frob(resf["newY"][0], reff["newY"]) < TOL_NEWY


def test_config_C3_full_size_and_properties(engine):
    """C3 (11,000 x 20,000, 400 sets x 3, k = 10): full structural oracle plus size-independent properties:
    multi-k snapshot == single-k run; per-gene shift equivariance; the centred control block of newY is
    orthogonal to ac (the normal equations of ruvIII.R:144)."""
    p = make_config("C3")
    m1 = p.meta["m1"]
    res = run_engine(engine, p, [4, 10], want_all=True)
    ref = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, 10,
                              apply_log=False, n_alpha_rows=10)
    assert O.rel_frob(res["newY"][1], ref["newY"][:m1]) < TOL_NEWY
    assert np.allclose(engine.eigenvalues(10), ref["eigenvalues"][:10], rtol=1e-10)
    single = run_engine(engine, p, [4], want_all=False)
    assert O.rel_frob(single["newY"][0], res["newY"][0]) < 1e-13
    # normal equations of ruvIII.R:144: (Ystand_c - W ac) ac' = 0
    ac = res["fullalpha"][:10][:, p.ctl]
    full_Yc = np.vstack([p.Y[:, p.ctl], p.replicate_data[:, p.ctl]])
    Ystand_c = full_Yc - full_Yc.mean(axis=0, keepdims=True)
    lhs = (Ystand_c - res["W"][1] @ ac) @ ac.T
    assert np.linalg.norm(lhs) < 1e-9 * np.linalg.norm(Ystand_c @ ac.T)
    # shift equivariance: adding a per-gene constant to every row (samples and PRPS) shifts newY by that constant
    shift = np.linspace(-1.0, 1.0, p.meta["n"])
    p2 = dataclasses.replace(p, assay=(p.Y + shift[None, :]).T, replicate_data=p.replicate_data + shift[None, :])
    res2 = run_engine(engine, p2, [10], want_all=False)
    assert O.rel_frob(res2["newY"][0] - shift[None, :], res["newY"][1]) < 1e-10


def test_user_fullalpha_row_counts(engine):
    """ruvIII.R:138,142: a supplied fullalpha skips residop / Gram / svd and alpha = fullalpha[1:min(k, nrow(fullalpha)), ] --
    whatever r is.  Fewer rows than k (W shrinks to nrow(fullalpha) columns; ADVICE r01: W came back partly uninitialised) and
    more rows than r = min(m - ncol(M), sum(ctl)) (used to be rejected)."""
    p = make_problem(m1=300, n=2000, groups=5, rep=3, n_ctl=150, k=6, seed=41)          # r = min(15 - 5, 150) = 10
    api.set_engine(engine)
    se = rv.SummarizedExperiment(assays={"logcounts": p.assay}, col_names=p.sample_names)
    rd = rv.ReplicateData(p.replicate_data, p.replicate_names)
    rng = np.random.default_rng(3)
    for rows, k in ((4, 6), (12, 12), (12, 7)):
        fa = rng.standard_normal((rows, 2000))
        out = rv.ruvIII(se, "logcounts", apply_log=False, replicate_data=rd, ctl=p.ctl, k=k, fullalpha=fa, return_info=True,
                        save_se_obj=False, verbose=False)
        ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, k, apply_log=False, fullalpha=fa)
        assert out["W"].shape == ref["W"].shape == (315, min(k, rows))
        assert O.rel_frob(out["W"], ref["W"]) < 1e-9
        assert O.rel_frob(out["newY"], ref["newY"]) < TOL_NEWY
    api.set_engine(None)
