"""Second, INDEPENDENT CPU oracle for the pieces of the RUV-III hot path that live in third-party R packages.

TEST INFRASTRUCTURE ONLY (only tests/ may import it).  PARITY UNPINNED against R -- like ruviii_oracle.py -- but the two
oracles share no arithmetic: where ruviii_oracle.py transcribes the R statements with the LAPACK drivers R calls, this file
reaches every quantity by a different numerical route, so a transcription slip in one cannot hide in code shared with the
other (VERDICT r01, next #1b).  tests/test_oracle.py requires the two to agree to <= 1e-10.

  quantity (reference)                          ruviii_oracle.py                      here
  -------------------------------------------   -----------------------------------   ----------------------------------------
  replicate.matrix(names)   ruvIII.R:101        np.unique one-hot                     dict of sorted levels, filled row by row
  residop(Y, M)             ruvIII.R:139        A - B solve(B'B) B'A  (normal eq.)    Y - Q Q'Y, Q from a Householder QR of M
  svd(Y0 Y0')$u             ruvIII.R:140        dgesdd of the m x m Gram              dgesvd (scipy, lapack_driver='gesvd') of Y0
                                                                                      itself: left vectors of Y0 = those of Y0Y0'
  runSVD(Y0, k)$u           fastruvIII.R:162    dgesdd of Y0                          eigh (dsyevr) of Y0 Y0', descending
  solve(ac ac')             ruvIII.R:144        explicit inverse (dgesv vs I)         least squares min ||ac' w - Ystand_c'|| by QR
  RUV1(Y, eta, ctl)         ruvIII.R:116        explicit inverse                      least squares by QR
  colMeans / centring       ruvIII.R:118-120    1 mu' subtraction                     projection onto the complement of 1
"""
import numpy as np
import scipy.linalg as sla


def _one_hot(names):
    # replicate.matrix orders its columns by the levels of factor(names), i.e. sorted (it matters for the row order of
    # apply.average.rep's output only)
    levels = {s: i for i, s in enumerate(sorted(set(names)))}
    M = np.zeros((len(names), len(levels)))
    for i, s in enumerate(names):
        M[i, levels[s]] = 1.0
    return M


def _residual_by_qr(Y, M):
    Q, _ = sla.qr(M, mode="economic")
    return Y - Q @ (Q.T @ Y)


def _tological(ctl, n):
    ctl = np.asarray(ctl)
    if ctl.dtype == bool and ctl.shape[0] == n:
        return ctl
    out = np.zeros(n, dtype=bool)
    out[ctl.astype(np.int64)] = True
    return out


def _design(eta, n, include_intercept):
    """t(ruv::design.matrix(eta)) as recalled from ruv 0.9.7.1: a scalar means intercept only; the column of ones is added
    when include.intercept and the constant vector is not already in the span of eta (residual^2 > 1e-8)."""
    e = np.asarray(eta, dtype=np.float64)
    if e.ndim == 0 or e.size == 1:
        return np.ones((1, n))
    e = np.atleast_2d(e)
    if e.shape[1] != n and e.shape[0] == n:
        e = e.T
    if include_intercept:
        ones = np.ones(n)
        coef = sla.lstsq(e.T, ones)[0]
        if np.sum((ones - e.T @ coef) ** 2) > 1e-8:
            e = np.vstack([ones[None, :], e])
    return e


def _ruv1(Y, eta, ctl, include_intercept):
    if eta is None:
        return Y
    e = _design(eta, Y.shape[1], include_intercept)
    # Y - Y_c eta_c' (eta_c eta_c')^-1 eta  ==  Y - B eta with B = argmin || eta_c' B' - Y_c' ||
    B = sla.lstsq(e[:, ctl].T, Y[:, ctl].T, lapack_driver="gelsy")[0].T
    return Y - B @ e


def _w(Ystand_c, ac):
    return sla.lstsq(ac.T, Ystand_c.T, lapack_driver="gelsy")[0].T


def ruvIII(assay, sample_names, replicate_data, replicate_names, ctl, k, apply_log=True, pseudo_count=1.0, eta=None,
           include_intercept=True, apply_average_rep=False):
    """ruvIII.R:86-149 by the routes listed in the module docstring.  dict(newY, fullalpha, W)."""
    expr = np.asarray(assay, dtype=np.float64)
    if apply_log:
        expr = np.log(expr + pseudo_count) / np.log(2.0)
    Y = np.concatenate([expr.T, np.asarray(replicate_data, dtype=np.float64)], axis=0)
    m, n = Y.shape
    M = _one_hot(list(sample_names) + list(replicate_names))
    ctl = _tological(ctl, n)
    Y = _ruv1(Y, eta, ctl, include_intercept)
    one = np.ones((m, 1)) / np.sqrt(m)
    Ystand = Y - one @ (one.T @ Y)
    W = fullalpha = None
    if M.shape[1] >= m or k == 0:
        newY = Y
    else:
        Y0 = _residual_by_qr(Y, M)
        r = min(m - M.shape[1], int(ctl.sum()))
        U = sla.svd(Y0, full_matrices=False, lapack_driver="gesvd")[0]
        fullalpha = U[:, :r].T @ Y
        alpha = fullalpha[:min(k, r)]
        W = _w(Ystand[:, ctl], alpha[:, ctl])
        newY = Y - W @ alpha
    if apply_average_rep:
        newY = (M / M.sum(axis=0, keepdims=True)).T @ newY
    return dict(newY=newY, fullalpha=fullalpha, W=W)


def fastruvIII(assay, replicate_data, replicate_names, ctl, k, apply_log=True, pseudo_count=1.0, eta=None,
               include_intercept=True):
    """fastruvIII.R:97-177 by the routes listed in the module docstring."""
    expr = np.asarray(assay, dtype=np.float64)
    if apply_log:
        expr = np.log(expr + pseudo_count) / np.log(2.0)
    Y = expr.T
    PS = np.asarray(replicate_data, dtype=np.float64)
    m1, n = Y.shape
    M = _one_hot(list(replicate_names))
    ctl = _tological(ctl, n)
    Y = _ruv1(Y, eta, ctl, include_intercept)
    PS = _ruv1(PS, eta, ctl, include_intercept)
    one = np.ones((m1, 1)) / np.sqrt(m1)
    Ystand = Y - one @ (one.T @ Y)
    W = fullalpha = None
    if M.shape[1] >= PS.shape[0] or k == 0:
        newY = Y
    else:
        Y0 = _residual_by_qr(PS, M)
        lam, V = sla.eigh(Y0 @ Y0.T)
        U = V[:, ::-1][:, :k]
        fullalpha = U.T @ PS
        W = _w(Ystand[:, ctl], fullalpha[:, ctl])
        newY = Y - W @ fullalpha
    return dict(newY=newY, fullalpha=fullalpha, W=W)
