"""K3 on spectra chosen to be hard for a block subspace iteration (VERDICT r01 "weak" item 3): a flat tail right below the
wanted pairs, slow geometric decay, a near-degenerate lambda_k ~ lambda_k+1, and the forced cuSOLVER fallback.

The replicate design is built so that the Gram the engine decomposes has a PRESCRIBED spectrum: every PRPS set has two
members a = c + d / sqrt(2), b = c - d / sqrt(2), so its Helmert contrast (kernels.cuh, K1) is exactly d, and the rows d form
Z = U diag(sqrt(lambda)) V' with random orthonormal U, V.  ruvIII.R:140 then decomposes Z Z' = U diag(lambda) U'.
Checked: eigenvalues against lambda, the eigenpair residual reconstructed from fullalpha = t(U) Z (ruvIII.R:140), newY
against the structural oracle, and that the device-side solver (not cuSOLVER) did the work."""
import os

import numpy as np
import pytest

import ruviii_b200 as rv
from oracle import ruviii_oracle as O
from ruviii_b200 import _cabi, api
from ruviii_b200.synth import Problem

pytestmark = pytest.mark.gpu

R, N, M1, NCTL, K = 256, 1536, 64, 300, 10


def spectrum_problem(lam, seed=5):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((R, R)))
    V, _ = np.linalg.qr(rng.standard_normal((N, R)))
    Z = (U * np.sqrt(lam)[None, :]) @ V.T                      # R x N, Z Z' = U diag(lam) U'
    centre = 8.0 + rng.standard_normal((R, N))
    PS = np.empty((2 * R, N))
    PS[0::2] = centre + Z / np.sqrt(2.0)
    PS[1::2] = centre - Z / np.sqrt(2.0)
    Y = 8.0 + rng.standard_normal((M1, N))
    ctl = np.zeros(N, dtype=bool)
    ctl[rng.choice(N, NCTL, replace=False)] = True
    names = ["S%04d" % i for i in range(M1)]
    rnames = ["set%04d" % (i // 2) for i in range(2 * R)]
    p = Problem(assay=np.ascontiguousarray(Y.T), sample_names=names, replicate_data=PS, replicate_names=rnames, ctl=ctl, k=K,
                meta=dict(m1=M1, n=N, n_ps=2 * R, groups=R, n_ctl=NCTL))
    return p, Z


def run(eng, p, ks):
    groups = api.group_codes(list(p.sample_names) + list(p.replicate_names))
    prm = _cabi.make_params(mode=_cabi.MODE_STACKED, want_ps_rows=True)
    return eng.normalise(prm, p.assay.T, p.replicate_data, groups, p.ctl, ks, want_fullalpha=True, want_W=True, want_ps=True)


def eigenpair_residual(Z, fullalpha, lam_top):
    """fullalpha_j = u_j' Z  =>  u_j = Z fullalpha_j' / lambda_j; returns max_j ||G u_j - lambda_j u_j|| / lambda_1 and the
    departure from orthonormality."""
    k = fullalpha.shape[0]
    lam = np.sum(fullalpha * fullalpha, axis=1)                # ||u' Z||^2 = lambda
    Uh = (Z @ fullalpha.T) / lam[None, :]
    G = Z @ Z.T
    res = np.linalg.norm(G @ Uh - Uh * lam[None, :], axis=0).max() / lam_top
    return res, np.abs(Uh.T @ Uh - np.eye(k)).max(), lam


j = np.arange(R)
SPECTRA = {
    "separated": 100.0 * 0.6 ** np.minimum(j, 40) * np.where(j < 40, 1.0, 0.5),
    "flat_tail": np.where(j < 12, 100.0 * 0.8 ** j, 100.0 * 0.8 ** 11 * 0.5),
    "slow_decay": 100.0 * 0.97 ** j,
    "near_degenerate_k": np.where(j == K, 100.0 * 0.97 ** (K - 1) / 1.01, 100.0 * 0.97 ** j),
    "power_law": 100.0 / (1.0 + j),
}


@pytest.mark.parametrize("name", list(SPECTRA))
def test_hard_spectra(name):
    lam = np.sort(SPECTRA[name])[::-1]
    p, Z = spectrum_problem(lam)
    with rv.Engine(devices=[0]) as eng:
        res = run(eng, p, [K])
        info = res["info"]
        assert info.eig_solver == 3, "the device-side subspace iteration must converge here, not fall back to cuSOLVER"
        assert 0 < info.eig_iterations <= 400
        ev = eng.eigenvalues(K)
    assert np.allclose(ev, lam[:K], rtol=1e-10)
    resid, orth, lam_hat = eigenpair_residual(Z, res["fullalpha"][:K], lam[0])
    assert resid < 1e-11 and orth < 1e-10
    assert np.allclose(lam_hat, lam[:K], rtol=1e-10)
    ref = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, K, apply_log=False)
    gap = (lam[K - 1] - lam[K]) / lam[0]
    tol = 1e-9 if gap > 1e-3 else 1e-7                          # newY is conditioned by 1 / (lambda_k - lambda_k+1)
    assert O.rel_frob(np.vstack([res["newY"][0], res["newY_ps"][0]]), ref["newY"]) < tol


def test_forced_fallback_to_cusolver():
    """RUVIII_EIG_MAXIT=2: the persistent solver gives up on the device, the flag travels back with the pass, and the
    engine reruns the tail with cuSOLVER syevdx -- same results."""
    lam = np.sort(SPECTRA["slow_decay"])[::-1]
    p, Z = spectrum_problem(lam)
    os.environ["RUVIII_EIG_MAXIT"] = "2"
    try:
        with rv.Engine(devices=[0]) as eng:
            res = run(eng, p, [K])
            assert res["info"].eig_solver == 2
            ev = eng.eigenvalues(K)
    finally:
        os.environ.pop("RUVIII_EIG_MAXIT", None)
    assert np.allclose(ev, lam[:K], rtol=1e-10)
    ref = O.ruvIII_structural(p.assay, p.sample_names, p.replicate_data, p.replicate_names, p.ctl, K, apply_log=False)
    assert O.rel_frob(np.vstack([res["newY"][0], res["newY_ps"][0]]), ref["newY"]) < 1e-9
    os.environ["RUVIII_EIG_MAXIT"] = "2"
    os.environ["RUVIII_EIG_IMPL"] = "3"
    try:
        with rv.Engine(devices=[0]) as eng:
            with pytest.raises(_cabi.RuviiiError) as ei:
                run(eng, p, [K])
            assert ei.value.code == _cabi.ERR_SOLVER
    finally:
        os.environ.pop("RUVIII_EIG_MAXIT", None)
        os.environ.pop("RUVIII_EIG_IMPL", None)


def test_persistent_matches_multikernel_solver():
    """RUVIII_EIG_PERSIST=0 selects the multi-kernel form of the same iteration (large orders use it): same eigenpairs."""
    lam = np.sort(SPECTRA["separated"])[::-1]
    p, Z = spectrum_problem(lam)
    outs = []
    for persist in ("1", "0"):
        os.environ["RUVIII_EIG_PERSIST"] = persist
        try:
            with rv.Engine(devices=[0]) as eng:
                res = run(eng, p, [K])
                assert res["info"].eig_solver == 3
                outs.append((res["newY"][0], eng.eigenvalues(K), res["fullalpha"][:K]))
        finally:
            os.environ.pop("RUVIII_EIG_PERSIST", None)
    assert np.allclose(outs[0][1], outs[1][1], rtol=1e-12)
    assert O.rel_frob(outs[0][0], outs[1][0]) < 1e-11
    # deterministic sign convention of both forms: the largest-magnitude component of every eigenvector is positive
    bad = {}
    for which, (_, ev, fa) in zip(("persistent", "multi-kernel"), outs):
        Uh = (Z @ fa.T) / ev[None, :]
        top = Uh[np.abs(Uh).argmax(axis=0), np.arange(K)]
        if not np.all(top > 0):
            bad[which] = np.nonzero(top <= 0)[0].tolist()
    assert not bad, "sign convention violated for columns %s" % bad
    assert O.rel_frob(outs[0][2], outs[1][2]) < 1e-8
