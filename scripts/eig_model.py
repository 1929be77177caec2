"""numpy model of the persistent eigensolver's iteration (ruviii_b200/csrc/eigen_persist.cu): same schedule, same degree
rule, same orthogonalisation (column-scaled Cholesky QR, second round under the same pivot test).  A development aid for
choosing the constants without a GPU: prints products / Rayleigh-Ritz steps / final residuals for a few spectra.

    python scripts/eig_model.py
"""
import numpy as np


def cholqr(Z, stats, accurate=True):
    """Column-scaled Cholesky QR; a second round below a pivot of 1e-4 (accurate: the block enters a Rayleigh-Ritz step) or
    1e-10 (the block only feeds the next product).  stats["minp"] = smallest pivot of the first round."""
    for rnd in range(2):
        S = Z.T @ Z
        d = 1.0 / np.sqrt(np.diag(S))
        A = 0.5 * (S + S.T) * d[:, None] * d[None, :]
        # right-looking Cholesky recording the pivots
        A = A.copy()
        n = A.shape[0]
        L = np.zeros_like(A)
        minp = 1.0
        for j in range(n):
            pj = A[j, j]
            if not pj > 1e-30:
                return None
            minp = min(minp, pj)
            L[j:, j] = A[j:, j] / np.sqrt(pj)
            A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
        Z = np.linalg.solve(L, (Z * d[None, :]).T).T
        if rnd == 0:
            stats["minp"] = minp
        if minp >= (1e-4 if accurate else 1e-10):
            break
        if rnd == 0:
            stats["chol2"] += 1
    return Z


def solve(G, kk, tol=1e-13, max_iter=150, first=6, period=3, cheb=True, B=32, seed=0, skip_minp=1e-3):
    R = G.shape[0]
    rng = np.random.default_rng(seed)
    stats = dict(products=0, rr=0, chol2=0, cheb=0, degs=[], qr=0, skipped=0)
    Q = rng.uniform(-1, 1, (R, B))                  # the kernel's start block is not orthogonalised either
    it, next_rr = 0, first
    skipped_last, last_minp = True, 0.0
    while it < max_iter:
        it += 1
        Z = G @ Q
        if it < next_rr:
            # one Cholesky QR for two products when the last factorisation found the block well conditioned; never twice
            # in a row, never right before a Rayleigh-Ritz step (eigen_persist.cu)
            if skip_minp > 0 and it >= 2 and not skipped_last and it + 1 < next_rr and last_minp >= skip_minp:
                Q = Z
                skipped_last = True
                stats["skipped"] += 1
                continue
            Q = cholqr(Z, stats, accurate=it + 1 >= next_rr)
            stats["qr"] += 1
            if Q is None:
                return None, stats
            last_minp = np.sqrt(stats["minp"]) if skipped_last and it > 1 else stats["minp"]
            skipped_last = False
            continue
        stats["rr"] += 1
        H = Q.T @ Z
        th, V = np.linalg.eigh(0.5 * (H + H.T))
        th, V = th[::-1], V[:, ::-1]
        X, Z = Q @ V, Z @ V
        res = np.linalg.norm(Z - X * th[None, :], axis=0)
        if np.all(res[:kk] <= tol * abs(th[0])):
            stats["products"] = it
            stats["resid"] = res[:kk].max() / abs(th[0])
            return (th, X), stats
        # continuation: filter rounds of degree <= dcap, each followed by Cholesky QR only, until the predicted gain on the
        # worst wanted pair reaches the tolerance; then one product + Rayleigh-Ritz
        bnd, th0 = max(th[B - 1], 1e-300), abs(th[0])
        tk, t0 = 2 * th[kk - 1] / bnd - 1, 2 * th0 / bnd - 1
        worst = res[:kk].max()
        plain = not cheb or not (tk > 1 + 1e-6)
        if plain:
            Q = cholqr(Z, stats)
            stats["qr"] += 1
            if Q is not None:
                last_minp, skipped_last = stats["minp"], False
            next_rr = it + period
        else:
            need = np.log(max(10 * worst / (tol * th0), 2.0))
            ak, a0 = np.arccosh(tk), np.arccosh(max(t0, 1 + 1e-6))
            dcap = max(1, int(np.floor((np.log(1e7 * (t0 + 1) / 2) + np.log(2)) / a0)))
            gain = 0.0
            Y0 = X
            e = c = 0.5 * bnd
            sig1 = e / (th0 - c)
            while gain < need and it < max_iter:
                left = need - gain
                d = int(min(dcap, max(1, np.ceil((left + np.log(2)) / ak)), 40, max_iter - it))
                stats["degs"].append(d)
                sig = sig1
                Yp, Y = None, Y0
                for i in range(1, d + 1):
                    GY = G @ Y - c * Y
                    sig_next = 1.0 / (2.0 / sig1 - sig)
                    Yn = (sig1 / e) * GY if i == 1 else (2 * sig_next / e) * GY - sig * sig_next * Yp
                    if i > 1:
                        sig = sig_next
                    Yp, Y = Y, Yn
                    it += 1
                    stats["cheb"] += 1
                gain += np.log(np.cosh(d * ak))
                Y0 = cholqr(Y, stats)
                if Y0 is None:
                    return None, stats
            Q = Y0
            skipped_last = True
            next_rr = it + 1
        if Q is None:
            return None, stats
    stats["products"] = it
    return None, stats


def spectrum_case(name, lam, R=800, kk=10, **kw):
    rng = np.random.default_rng(1)
    U, _ = np.linalg.qr(rng.standard_normal((R, R)))
    G = (U * lam[None, :]) @ U.T
    G = 0.5 * (G + G.T)
    out, st = solve(G, kk, **kw)
    ok = out is not None
    err = None
    if ok:
        err = np.abs(out[0][:kk] - lam[:kk]).max() / lam[0]
    print("%-28s converged=%s products=%d qr=%d skipped=%d rr=%d cheb=%d degs=%s chol2=%d resid=%s eval_err=%s" % (
        name, ok, st["products"], st["qr"], st["skipped"], st["rr"], st["cheb"], st["degs"], st["chol2"], st.get("resid"), err))


if __name__ == "__main__":
    R = 800
    j = np.arange(R)
    planted = np.where(j < 24, 2.25 * 0.7225 ** j * 3000, 1.0) * (1 + 0.3 * np.random.default_rng(2).random(R)) ** np.where(j < 24, 0, 1)
    spectrum_case("C3-like (planted 24)", np.sort(planted)[::-1])
    spectrum_case("C3-like, QR after every product", np.sort(planted)[::-1], skip_minp=0.0)
    spectrum_case("C3-like, no chebyshev", np.sort(planted)[::-1], cheb=False)
    flat = np.where(j < 12, 100 * 0.8 ** j, 100 * 0.8 ** 11 * 0.5)
    spectrum_case("flat tail (ratio 0.5)", flat)
    spectrum_case("flat tail, no chebyshev", flat, cheb=False)
    slow = 100 * 0.97 ** j
    spectrum_case("slow decay 0.97^j", slow)
    spectrum_case("slow decay, no chebyshev", slow, cheb=False)
    close = slow.copy()
    close[10] = close[9] / 1.01
    spectrum_case("lambda_k / lambda_k+1 = 1.01", close)
    vslow = 100 * 0.995 ** j
    spectrum_case("very slow 0.995^j", vslow)
    spectrum_case("very slow, k=20", vslow, kk=20, max_iter=45)
    power = 100.0 / (1 + j) ** 1.0
    spectrum_case("power law 1/j", power)
    spectrum_case("power law 1/j, k=20", power, kk=20, max_iter=45)
