"""numpy model of the systolic ("chair" order) Jacobi of ruviii_b200/csrc/eigen_persist.cu on the Rayleigh-Ritz matrix of a
C3-like problem: prints the largest relative off-diagonal entry rotated in every sweep for the initial chair assignments
RUVIII_EIG_JORDER = 0 / 1 / 2, i.e. how many sweeps each needs and what the stopping threshold may be.  Development aid (CPU).

    python scripts/jacobi_model.py
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ruviii_b200.synth import make_problem

PB, H = 32, 16


def helmert_rows(PS, rep):
    out = []
    for g in range(PS.shape[0] // rep):
        y = PS[g * rep:(g + 1) * rep]
        for j in range(1, rep):
            out.append(np.sqrt(j / (j + 1.0)) * (y[:j].mean(axis=0) - y[j]))
    return np.array(out)


def rr_matrix(G, products=6, seed=0):
    rng = np.random.default_rng(seed)
    Q = np.linalg.qr(G @ rng.uniform(-1, 1, (G.shape[0], PB)))[0]
    for _ in range(products - 2):
        Q = np.linalg.qr(G @ Q)[0]
    Hm = Q.T @ (G @ Q)
    return 0.5 * (Hm + Hm.T)


def nxt(c):                                    # chair c (0..15 = T, 16..31 = B) -> the chair its value moves to
    s, i = divmod(c, H)
    if s == 0:
        return 0 if i == 0 else (H + H - 1 if i == H - 1 else i + 1)
    return 1 if i == 0 else H + i - 1


def jacobi(Hm, order):
    idx = [(2 * (c % H) + c // H) if order == 0 else ((PB - 1 - c % H) if c // H else c % H) if order == 1 else c for c in range(PB)]
    A = Hm[np.ix_(idx, idx)].copy()
    perm = np.empty(PB, dtype=int)
    for c in range(PB):
        perm[nxt(c)] = c                       # new[nxt(c)] = old[c]
    eps = []
    for sweep in range(12):
        emax = 0.0
        for rnd in range(PB - 1):
            J = np.eye(PB)
            for i in range(H):
                p, q = i, H + i
                app, aqq, apq = A[p, p], A[q, q], A[p, q]
                if apq * apq > 1e-33 * abs(app * aqq):
                    emax = max(emax, abs(apq) / np.sqrt(abs(app * aqq)))
                    zeta, beta = aqq - app, 2.0 * apq
                    rr = np.hypot(zeta, beta)
                    c2 = 0.5 + 0.5 * abs(zeta) / rr
                    c = np.sqrt(c2)
                    s = (0.5 if zeta >= 0 else -0.5) * beta / rr / c
                    J[p, p] = c; J[q, q] = c; J[p, q] = s; J[q, p] = -s
            A = J.T @ A @ J
            A = A[np.ix_(perm, perm)]
        eps.append(emax)
        if emax < 1e-10:
            break
    return eps, np.sort(np.diag(A))[::-1]


if __name__ == "__main__":
    for name, kw in (("C3-like (n = 4000)", dict(m1=2000, n=4000, groups=400, rep=3, n_ctl=500, k=10)),
                     ("small (n = 1500)", dict(m1=600, n=1500, groups=300, rep=3, n_ctl=200, k=5))):
        p = make_problem(**kw)
        Z = helmert_rows(p.replicate_data, 3)
        G = Z @ Z.T
        Hm = rr_matrix(G)
        ref = np.sort(np.linalg.eigvalsh(Hm))[::-1]
        print(name, "R =", G.shape[0], " top Ritz values", np.round(ref[:12], 1))
        for order in (0, 1, 2):
            eps, th = jacobi(Hm, order)
            print("  order %d: sweeps %d, max relative off-diagonal rotated per sweep:" % (order, len(eps)),
                  " ".join("%.1e" % e for e in eps), " eig err %.1e" % (np.abs(th - ref).max() / ref[0]))
