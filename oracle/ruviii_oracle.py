"""CPU oracle for the RUV-III hot path of RMolania/RUVIIIPRPS.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The shipped path
(``ruviii_b200`` -> ``libruviii_b200.so``) never imports, links or executes anything in ``oracle/``.

PARITY UNPINNED.  The reference is a pure-R package with no tests, no golden vectors and no
example data, R itself is not installed in the build container, and part of the
arithmetic lives in third-party packages that are not in the tree and are not version-pinned
(``DESCRIPTION:14,33``):

  * ``ruv`` (CRAN, latest known 0.9.7.1): ``replicate.matrix``, ``residop``, ``RUV1``
  * ``BiocSingular`` (Bioconductor): ``runSVD`` / ``bsparam``  (default ``ExactParam`` = truncated exact SVD)
  * base R: ``svd`` -> LAPACK dgesdd, ``solve`` -> LAPACK dgesv, ``%*%`` -> BLAS dgemm

So this module restates the R code line by line with numpy's bindings of the *same LAPACK drivers*
(``numpy.linalg.svd`` = dgesdd, ``numpy.linalg.inv``/``solve`` = dgesv, ``@`` = dgemm) and uses the
documented semantics for the third-party pieces.  It was never compared with output of a real R
session; every parity claim made against it is therefore capped at "matches the restatement".

Two flavours:

  * ``*_literal``   follow the R source statement by statement (dense 0/1 ``M``, 5-GEMM residop, full
                    ``Y0 %*% t(Y0)``, full ``svd``).  Cost grows like m^3; used for small/medium shapes
                    and as the CPU baseline that stands in for R.
  * ``*_structural`` use the algebra of SURVEY.md appendix A (group-mean residop, eigen-decomposition of the
                    non-singleton block only).  Used to check large shapes.  ``tests/test_oracle.py`` pins the
                    two flavours against each other.

Conventions: the assay is ``genes x samples`` (n x m1) as in a SummarizedExperiment; ``replicate_data`` is
``pseudo-samples x genes`` (n_ps x n) with one name per row (``normalise.R:143``); ``ctl`` is a boolean
vector of length n or an index vector (0-based here; R's is 1-based).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "tological", "replicate_matrix", "residop", "RUV1",
    "ruvIII_literal", "fastruvIII_literal", "ruvIIIMultipleK_literal",
    "ruvIII_structural", "fastruvIII_structural", "ruvIIIMultipleK_structural",
    "group_ids_from_names", "rel_frob", "align_rows_sign",
]


# --------------------------------------------------------------------------------------------
# helpers shared by both flavours
# --------------------------------------------------------------------------------------------

def tological(ctl, n):
    """``tological`` (ruvIII.R:80-84, fastruvIII.R:83-87): index or logical -> logical[n]."""
    ctl = np.asarray(ctl)
    if ctl.dtype == np.bool_:
        if ctl.shape[0] != n:
            raise ValueError("logical ctl must have length n")
        return ctl.copy()
    ctl2 = np.zeros(n, dtype=bool)
    ctl2[ctl.astype(np.int64)] = True
    return ctl2


def replicate_matrix(names):
    """``ruv::replicate.matrix(a)`` for a character vector (ruvIII.R:101, fastruvIII.R:117-118).

    Documented semantics: one row per element, one column per distinct value (``factor(a)`` levels,
    i.e. sorted), entry 1 where the element equals the level.  The column order does not affect any
    downstream quantity (only ``M (M'M)^-1 M'`` and ``ncol(M)`` are used).
    """
    names = np.asarray(names)
    levels, inv = np.unique(names, return_inverse=True)
    M = np.zeros((names.shape[0], levels.shape[0]), dtype=np.float64)
    M[np.arange(names.shape[0]), inv] = 1.0
    return M


def group_ids_from_names(names):
    """0-based group id per row, numbered in order of first appearance (not part of R; a helper for tests)."""
    seen = {}
    out = np.empty(len(names), dtype=np.int32)
    for i, s in enumerate(names):
        out[i] = seen.setdefault(s, len(seen))
    return out


def residop(A, B):
    """``fast.residop`` (fastruvIII.R:88-95), the in-tree statement of ``ruv::residop`` (ruvIII.R:139).

    A - B (B'B)^-1 B' A, evaluated in the same order as the R code.
    """
    tBB = B.T @ B                       # fastruvIII.R:89
    tBB_inv = np.linalg.inv(tBB)        # :90  solve(tBB) -> dgesv against the identity
    BtBB_inv = B @ tBB_inv              # :91
    tBA = B.T @ A                       # :92
    BtBB_inv_tBA = BtBB_inv @ tBA       # :93
    return A - BtBB_inv_tBA             # :94


def design_matrix_t(eta, n, include_intercept=True):
    """``t(ruv::design.matrix(eta, include.intercept))`` as ``ruv::RUV1`` uses it: q x n, one column per gene.

    Restated from memory of ruv 0.9.7.1 (the package source is not in the container): ``RUV1`` turns a single number
    into ``matrix(1, n, 1)`` (intercept only); ``design.matrix`` prepends the column of ones when ``include.intercept``
    and the constant vector is not already in the column span of the design
    (``sum(residop(matrix(1, n, 1), A)^2) > 1e-8``).  A gene-wise matrix may be given n x q (R) or q x n.
    """
    e = np.asarray(eta, dtype=np.float64)
    if e.ndim == 0 or e.size == 1:
        return np.ones((1, n))
    e = np.atleast_2d(e)
    if e.shape[1] != n and e.shape[0] == n:
        e = e.T
    if e.shape[1] != n:
        raise ValueError("eta must have one row per gene (n x q)")
    if include_intercept:
        A = e.T                                                   # n x q, the design as R holds it
        ones = np.ones((n, 1))
        if float(np.sum(residop(ones, A) ** 2)) > 1e-8:
            e = np.vstack([np.ones((1, n)), e])
    return e


def RUV1(Y, eta, ctl, include_intercept=True):
    """``ruv::RUV1`` (ruvIII.R:116, fastruvIII.R:133-134).  Identity for ``eta = NULL`` (the default,
    normalise.R:81).  The ``eta`` branch restates ruv 0.9.7.1 from memory of its published source
    (``Y - Y_c eta_c' (eta_c eta_c')^-1 eta`` with ``eta = t(design.matrix(eta))``, see :func:`design_matrix_t`);
    the package source is not in the container, so that branch is doubly unpinned.
    """
    if eta is None:
        return Y
    eta = design_matrix_t(eta, Y.shape[1], include_intercept)
    Yc = Y[:, ctl]
    etac = eta[:, ctl]
    return Y - Yc @ etac.T @ np.linalg.inv(etac @ etac.T) @ eta


def rel_frob(a, b):
    """||a-b||_F / ||b||_F (the north-star's comparison metric; b is the reference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def align_rows_sign(a, ref):
    """Flip the sign of each row of ``a`` so that <a_i, ref_i> >= 0 (singular vectors are sign-free)."""
    a = np.array(a, dtype=np.float64, copy=True)
    s = np.sign(np.einsum("ij,ij->i", a, ref))
    s[s == 0] = 1.0
    return a * s[:, None]


def _log_assay(assay, apply_log, pseudo_count):
    # ruvIII.R:88-92 / fastruvIII.R:99-103
    assay = np.asarray(assay, dtype=np.float64)
    return np.log2(assay + pseudo_count) if apply_log else assay


# --------------------------------------------------------------------------------------------
# literal restatement
# --------------------------------------------------------------------------------------------

def ruvIII_literal(assay, sample_names, replicate_data, replicate_names, ctl, k,
                   apply_log=True, pseudo_count=1.0, eta=None, include_intercept=True,
                   apply_average_rep=False, fullalpha=None):
    """Statement-by-statement restatement of ``ruvIII`` (ruvIII.R:86-149).

    Returns ``dict(newY, fullalpha, W, m1)`` exactly as the ``return.info & !save.se.obj`` branch does
    (ruvIII.R:193-197): ``newY`` is the full (m x n) matrix including the PRPS rows; the assay written
    by ``save.se.obj`` is ``newY[:m1].T`` (ruvIII.R:158).
    """
    if k is None:
        raise ValueError("k cannot be 0. This means no adjustment will be made.")   # ruvIII.R:62-63
    expr = _log_assay(assay, apply_log, pseudo_count)                                 # :88-92
    replicate_data = np.asarray(replicate_data, dtype=np.float64)
    Y = np.vstack([expr.T, replicate_data])                                           # :93
    m, n = Y.shape                                                                    # :99-100
    m1 = expr.shape[1]
    M = replicate_matrix(list(sample_names) + list(replicate_names))                  # :101
    ctl = tological(ctl, n)                                                           # :102
    Y = RUV1(Y, eta, ctl, include_intercept)                                          # :116
    mu = Y.mean(axis=0)                                                               # :118
    mu_mat = np.ones((m, 1)) @ mu[None, :]                                            # :119
    Y_stand = Y - mu_mat                                                              # :120
    W = None
    if M.shape[1] >= m:                                                               # :128-129
        newY = Y
    elif k == 0:                                                                      # :134-136
        newY = Y
        fullalpha = None
    else:
        if fullalpha is None:                                                         # :138
            Y0 = residop(Y, M)                                                        # :139
            r = min(m - M.shape[1], int(ctl.sum()))
            U = np.linalg.svd(Y0 @ Y0.T)[0]                                           # :140 svd()$u (dgesdd)
            fullalpha = U[:, :r].T @ Y                                                # :140
        alpha = fullalpha[:min(k, fullalpha.shape[0]), :]                             # :142
        ac = alpha[:, ctl]                                                            # :143
        W = Y_stand[:, ctl] @ ac.T @ np.linalg.inv(ac @ ac.T)                         # :144
        newY = Y - W @ alpha                                                          # :145
    if apply_average_rep:                                                             # :148-149
        newY = ((1.0 / M.sum(axis=0))[:, None] * M.T) @ newY
    return dict(newY=newY, fullalpha=fullalpha, W=W, m1=m1)


def fastruvIII_literal(assay, replicate_data, replicate_names, ctl, k,
                       apply_log=True, pseudo_count=1.0, eta=None, include_intercept=True,
                       apply_average_rep=False, fullalpha=None):
    """Statement-by-statement restatement of ``fastruvIII`` (fastruvIII.R:97-180).

    ``runSVD(x, k, BSPARAM = bsparam())`` with the default ``ExactParam`` is base ``svd(x, nu = k, nv = k)``
    (BiocSingular documentation), i.e. the leading k left singular vectors from dgesdd.
    """
    if k is None:
        raise ValueError("k cannot be 0. This means no adjustment will be made.")   # fastruvIII.R:65-66
    names, counts = np.unique(np.asarray(replicate_names), return_counts=True)
    if counts.min() == 1:                                                             # :67-68
        raise ValueError("There are only replicated samples of a single sample in the replicate.data")
    expr = _log_assay(assay, apply_log, pseudo_count)                                 # :99-103
    Y = expr.T                                                                        # :104
    replicate_data = np.asarray(replicate_data, dtype=np.float64)
    m = replicate_data.shape[0]                                                       # :112
    m1, n = Y.shape                                                                   # :114,116
    M = replicate_matrix(replicate_names)                                             # :117-118
    ctl = tological(ctl, n)                                                           # :119
    Y = RUV1(Y, eta, ctl, include_intercept)                                          # :133
    replicate_data = RUV1(replicate_data, eta, ctl, include_intercept)                # :134
    mu = Y.mean(axis=0)                                                               # :136
    Y_stand = Y - np.ones((m1, 1)) @ mu[None, :]                                      # :137-138
    W = None
    if M.shape[1] >= m:                                                               # :146-147
        newY = Y
    elif k == 0:                                                                      # :152-154
        newY = Y
        fullalpha = None
    else:
        if fullalpha is None:
            Y0 = residop(replicate_data, M)                                           # :157
            eigVec = np.linalg.svd(Y0, full_matrices=False)[0][:, :k]                 # :162-168
            fullalpha = eigVec.T @ replicate_data                                     # :171
        alpha = fullalpha[:min(k, fullalpha.shape[0]), :]                             # :173
        ac = alpha[:, ctl]                                                            # :174
        W = Y_stand[:, ctl] @ ac.T @ np.linalg.inv(ac @ ac.T)                         # :175
        newY = Y - W @ alpha                                                          # :176
    if apply_average_rep:                                                             # :179-180 (as written it
        newY = ((1.0 / M.sum(axis=0))[:, None] * M.T) @ newY                          #  only conforms if m1 == m)
    return dict(newY=newY, fullalpha=fullalpha, W=W, m1=m1)


def ruvIIIMultipleK_literal(assay, sample_names, replicate_data, replicate_names, ctl, ks, **kw):
    """``ruvIIIMultipleK`` (ruvIIIMultipleK.R:75-94): one complete ``ruvIII`` per k, nothing shared
    between iterations (the user's ``fullalpha``, default NULL, is forwarded unchanged, :87)."""
    if ks is None:
        raise ValueError("k cannot be 0. This means no adjustment will be made.")   # ruvIIIMultipleK.R:57-58
    return {int(k): ruvIII_literal(assay, sample_names, replicate_data, replicate_names, ctl, int(k), **kw)
            for k in ks}


# --------------------------------------------------------------------------------------------
# structural restatement: same quantities, no dense M, no m x m SVD
# --------------------------------------------------------------------------------------------

def _group_mean_residual(Y, gid):
    """Rows of non-singleton groups, their residual after subtracting the group mean, and p.

    Equals ``residop(Y, replicate.matrix(names))`` restricted to its nonzero rows (appendix A-1).
    """
    gid = np.asarray(gid)
    p = int(gid.max()) + 1
    cnt = np.bincount(gid, minlength=p)
    active = np.nonzero(cnt[gid] > 1)[0]
    ga = gid[active]
    order = np.argsort(ga, kind="stable")
    active = active[order]
    ga = ga[order]
    Ya = Y[active]
    uniq, start = np.unique(ga, return_index=True)
    sums = np.add.reduceat(Ya, start, axis=0)
    sizes = np.diff(np.append(start, len(ga)))
    means = sums / sizes[:, None]
    Y0a = Ya - np.repeat(means, sizes, axis=0)
    return active, Y0a, p


def _top_left_vectors(Y0a, r):
    """Leading r left singular vectors of Y0a via the smaller Gram side (appendix A-3, A-5)."""
    R, n = Y0a.shape
    if R <= n:
        lam, U = np.linalg.eigh(Y0a @ Y0a.T)
        idx = np.argsort(lam)[::-1][:r]
        return U[:, idx], lam[idx]
    lam, V = np.linalg.eigh(Y0a.T @ Y0a)
    idx = np.argsort(lam)[::-1][:r]
    lam_r = np.maximum(lam[idx], 0.0)
    U = (Y0a @ V[:, idx]) / np.sqrt(lam_r)[None, :]
    return U, lam_r


def _w_newy(Y, Y_mu_rows, alpha, ctl):
    """W and newY from alpha with the centring trick (appendix A-6): mean of A over the mu rows."""
    ac = alpha[:, ctl]
    A = Y[:, ctl] @ ac.T
    A = A - A[:Y_mu_rows].mean(axis=0, keepdims=True)
    W = np.linalg.solve(ac @ ac.T, A.T).T
    return W, Y - W @ alpha


def ruvIII_structural(assay, sample_names, replicate_data, replicate_names, ctl, k,
                      apply_log=True, pseudo_count=1.0, n_alpha_rows=None):
    """Same outputs as :func:`ruvIII_literal` (``eta = NULL``, no averaging) computed structurally.

    ``n_alpha_rows`` limits how many rows of ``fullalpha`` are produced (default: all r); ``newY`` and ``W``
    need only the first k.
    """
    expr = _log_assay(assay, apply_log, pseudo_count)
    Y = np.vstack([expr.T, np.asarray(replicate_data, dtype=np.float64)])
    m, n = Y.shape
    m1 = expr.shape[1]
    gid = group_ids_from_names(list(sample_names) + list(replicate_names))
    ctl = tological(ctl, n)
    active, Y0a, p = _group_mean_residual(Y, gid)
    if p >= m or k == 0:
        return dict(newY=Y, fullalpha=None, W=None, m1=m1)
    r = min(m - p, int(ctl.sum()))
    if n_alpha_rows is not None:
        r = min(r, max(int(n_alpha_rows), k))
    Ua, lam = _top_left_vectors(Y0a, r)
    fullalpha = Ua.T @ Y[active]
    alpha = fullalpha[:min(k, r)]
    W, newY = _w_newy(Y, m, alpha, ctl)
    return dict(newY=newY, fullalpha=fullalpha, W=W, m1=m1, eigenvalues=lam)


def fastruvIII_structural(assay, replicate_data, replicate_names, ctl, k,
                          apply_log=True, pseudo_count=1.0):
    """Same outputs as :func:`fastruvIII_literal` (``eta = NULL``, no averaging) computed structurally."""
    expr = _log_assay(assay, apply_log, pseudo_count)
    Y = expr.T
    m1, n = Y.shape
    PS = np.asarray(replicate_data, dtype=np.float64)
    gid = group_ids_from_names(list(replicate_names))
    ctl = tological(ctl, n)
    active, Y0a, p = _group_mean_residual(PS, gid)
    if p >= PS.shape[0] or k == 0:
        return dict(newY=Y, fullalpha=None, W=None, m1=m1)
    Ua, lam = _top_left_vectors(Y0a, k)
    fullalpha = Ua.T @ PS[active]
    alpha = fullalpha[:min(k, fullalpha.shape[0])]
    W, newY = _w_newy(Y, m1, alpha, ctl)
    return dict(newY=newY, fullalpha=fullalpha, W=W, m1=m1, eigenvalues=lam)


def ruvIIIMultipleK_structural(assay, sample_names, replicate_data, replicate_names, ctl, ks,
                               apply_log=True, pseudo_count=1.0):
    """All k from one decomposition (what the engine does); equals the per-k loop of the reference."""
    ks = [int(k) for k in ks]
    kmax = max(ks)
    base = ruvIII_structural(assay, sample_names, replicate_data, replicate_names, ctl, kmax,
                             apply_log=apply_log, pseudo_count=pseudo_count, n_alpha_rows=kmax)
    expr = _log_assay(assay, apply_log, pseudo_count)
    Y = np.vstack([expr.T, np.asarray(replicate_data, dtype=np.float64)])
    ctlb = tological(ctl, Y.shape[1])
    out = {}
    for k in ks:
        if k == 0 or base["fullalpha"] is None:
            out[k] = dict(newY=Y, fullalpha=base["fullalpha"], W=None, m1=base["m1"])
            continue
        alpha = base["fullalpha"][:min(k, base["fullalpha"].shape[0])]
        W, newY = _w_newy(Y, Y.shape[0], alpha, ctlb)
        out[k] = dict(newY=newY, fullalpha=base["fullalpha"], W=W, m1=base["m1"])
    return out


def ruvIII_structural_rows(assay, sample_names, replicate_data, replicate_names, ctl, ks, rows,
                           apply_log=True, pseudo_count=1.0):
    """``newY[rows, ]`` (sample rows only) for every k in ``ks`` -- the arithmetic of :func:`ruvIIIMultipleK_structural`
    without materialising any m x n result, so that a full-size run (11,000 x 20,000, k = 1..20) can be spot-checked in
    seconds: one decomposition of the residual of the replicated rows, then per k
    ``W = Ystand_c ac' (ac ac')^-1`` (ruvIII.R:144) over ALL m rows and ``newY[rows] = Y[rows] - W[rows] alpha`` (:145).
    Returns ``(dict k -> len(rows) x n, eigenvalues, fullalpha[:max(ks)])``."""
    expr = _log_assay(assay, apply_log, pseudo_count)
    Ys = expr.T
    PS = np.asarray(replicate_data, dtype=np.float64)
    m1, n = Ys.shape
    m = m1 + PS.shape[0]
    gid = group_ids_from_names(list(sample_names) + list(replicate_names))
    ctl = tological(ctl, n)
    p = int(gid.max()) + 1
    cnt = np.bincount(gid, minlength=p)
    act = np.nonzero(cnt[gid] > 1)[0]
    rows = np.asarray(rows)
    ks = [int(k) for k in ks]
    kmax = max(ks)
    out = {}
    if p >= m or kmax == 0 or act.size == 0:
        return {k: np.array(Ys[rows]) for k in ks}, None, None
    Ya = np.vstack([Ys[act[act < m1]], PS[act[act >= m1] - m1]])
    ga = np.unique(gid[act], return_inverse=True)[1]
    order, Y0a, _ = _group_mean_residual(Ya, ga)
    r = min(m - p, int(ctl.sum()))
    Ua, lam = _top_left_vectors(Y0a, min(r, kmax))
    fullalpha = Ua.T @ Ya[order]
    Yc = np.vstack([Ys[:, ctl], PS[:, ctl]])
    for k in ks:
        if k == 0:
            out[k] = np.array(Ys[rows])
            continue
        alpha = fullalpha[:min(k, fullalpha.shape[0])]
        ac = alpha[:, ctl]
        A = Yc @ ac.T
        A = A - A.mean(axis=0, keepdims=True)
        W = np.linalg.solve(ac @ ac.T, A.T).T
        out[k] = Ys[rows] - W[rows] @ alpha
    return out, lam, fullalpha
