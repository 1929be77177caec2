"""CPU ORACLE (test infrastructure, never shipped, never imported by ruviii_b200/) for the statistics of
``supervisedFindNCG`` (R/supervisedFindNCG.R:179-191, 214-229, 270-285, 302-316, 342-357).

``matrixTests::row_oneway_equalvar`` and ``Rfast::correls`` are not vendored in the reference and R is not installed:
PARITY UNPINNED.  They are restated from their documentation -- the classical equal-variance one-way ANOVA F statistic
(group means and within-group sums of squares, two passes, as matrixTests computes it) and the Pearson correlation."""
import numpy as np


def row_oneway_equalvar_statistic(x, g):
    """x: genes x samples, g: one label per sample.  Returns F per gene."""
    x = np.asarray(x, dtype=np.float64)
    g = np.asarray(g)
    levels = np.unique(g)
    n_tot = x.shape[1]
    grand = x.mean(axis=1)
    ssb = np.zeros(x.shape[0])
    ssw = np.zeros(x.shape[0])
    for lev in levels:
        xs = x[:, g == lev]
        m = xs.mean(axis=1)
        ssb += xs.shape[1] * (m - grand) ** 2
        ssw += ((xs - m[:, None]) ** 2).sum(axis=1)
    return (ssb / (levels.size - 1)) / (ssw / (n_tot - levels.size))


def correls_pearson(y, x):
    """Rfast::correls(y, x, type = "pearson")[, "correlation"]: x is samples x genes, y one value per sample."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xc = x - x.mean(axis=0, keepdims=True)
    yc = y - y.mean()
    return (xc * yc[:, None]).sum(axis=0) / np.sqrt((xc ** 2).sum(axis=0) * (yc ** 2).sum())


def correls_spearman(y, x):
    """Rfast::correls(y, x, type = "spearman")[, "correlation"]: Pearson's r of the average ranks (scipy.stats.rankdata, an
    implementation independent of rank_average below)."""
    from scipy.stats import rankdata
    x = np.asarray(x, dtype=np.float64)
    return correls_pearson(rankdata(np.asarray(y, dtype=np.float64)), rankdata(x, axis=0))


def rank_average(x):
    """base::rank, ties.method = "average" (via a double argsort with tie groups averaged)."""
    x = np.asarray(x, dtype=np.float64)
    uniq, inv, counts = np.unique(x, return_inverse=True, return_counts=True)
    ends = np.cumsum(counts)
    starts = ends - counts
    return (0.5 * (starts + ends - 1) + 1.0)[inv]


def lm_residuals(expr, col_data, variables):
    """t(lm(t(expr) ~ v1 + v2 + ...)$residuals), supervisedFindNCG.R:111-133: least squares by SVD (numpy.linalg.lstsq: the
    minimum-norm solution, so a rank-deficient model matrix gives the residuals lm's pivoted QR gives).  expr: genes x samples."""
    m = expr.shape[1]
    cols = [np.ones(m)]
    for v in variables:
        c = col_data[v]
        if all(isinstance(t, str) for t in c):
            c = np.asarray(c)
            cols += [(c == lev).astype(np.float64) for lev in sorted(set(c.tolist()))[1:]]
        else:
            cols.append(np.asarray(c, dtype=np.float64))
    D = np.stack(cols, axis=1)
    beta = np.linalg.lstsq(D, expr.T, rcond=1e-10)[0]
    return (expr.T - D @ beta).T


def supervised_find_ncg(assay, col_data, bio_variables, uv_variables, no_ncg, apply_log=True, pseudo_count=1.0, corr_method="pearson",
                        regress_out_uv=False, regress_out_bio=False):
    """The flow of supervisedFindNCG.R:104-357 without apply.normalization (see ruviii_b200/ncg.py for the quirks)."""
    correls_pearson = globals()["correls_spearman" if corr_method == "spearman" else "correls_pearson"]
    expr = np.log2(np.asarray(assay, dtype=np.float64) + pseudo_count) if apply_log else np.asarray(assay, dtype=np.float64)
    uv_data = bio_data = expr
    if regress_out_uv:                                                        # :111-121 -> used by the biology tests (:248-254)
        bio_data = lm_residuals(expr, col_data, uv_variables)
    elif regress_out_bio:                                                     # :122-133 -> used by the UV tests (:163-167)
        uv_data = lm_residuals(expr, col_data, bio_variables)
    is_cat = lambda v: all(isinstance(t, str) for t in col_data[v])
    ranks = []
    for v in [u for u in uv_variables if is_cat(u)]:
        g = np.asarray(col_data[v])
        lev, cnt = np.unique(g, return_counts=True)
        keep = ~np.isin(g, lev[cnt < 3])
        ranks.append(rank_average(row_oneway_equalvar_statistic(uv_data[:, keep], g[keep])))
    for v in [u for u in uv_variables if not is_cat(u)]:
        ranks.append(rank_average(np.abs(np.round(correls_pearson(col_data[v], uv_data.T), 2))))
    for v in [b for b in bio_variables if not is_cat(b)]:
        ranks.append(rank_average(-np.abs(np.round(correls_pearson(col_data[v], bio_data.T), 2))))
    for v in [b for b in bio_variables if is_cat(b)]:
        ranks.append(rank_average(-row_oneway_equalvar_statistic(bio_data, np.asarray(col_data[v]))))
    prod = np.prod(np.stack(ranks, axis=1), axis=1)
    return rank_average(-prod) < no_ncg + 1
