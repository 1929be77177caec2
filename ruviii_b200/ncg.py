"""Negative-control-gene selection -- the other step immediately before the RUV-III path.

Host side of ``supervisedFindNCG`` (R/supervisedFindNCG.R:44-461) with ``apply.normalization = FALSE``, for
``corr.method`` "pearson" (default) and "spearman": the same arguments, the same sequence of tests, ranks and rank
product.  ``regress.out.uv.variables`` / ``regress.out.bio.variables`` (:111-133, which the reference itself marks "NEED to
BE tEStED") are ``ruviii_regress_out``: the model matrix of the ``lm`` formula is built here (indexing only), the fit and
the residuals are device work (K12 + K4 + K6) and the statistics then read the residuals in place (``slot = -2``).  The per-gene statistics -- one-way ANOVA F across every categorical variable, Pearson correlation with every
continuous one -- are the streaming kernel K11 through ``ruviii_gene_stats``; Spearman's rho is K11b
(``ruviii_gene_spearman``: a shared-memory sort of every gene's samples);
``rank()``, the ``round(., 2)`` the reference applies to the correlations, and the rank product are O(genes) host
bookkeeping.  Plotting (:359-425) is not reproduced.

Reference quirks kept or resolved (stated here because they change which genes come out):
  * :183-187  levels of a categorical UV variable with fewer than three samples are dropped from the data, but the
    grouping vector handed to row_oneway_equalvar is not subset (an error in R whenever a level is dropped).  Here
    the samples of such levels are left out of both.
  * :111-133  ``if (regress.out.uv.variables) ... else if (regress.out.bio.variables)``: with both TRUE only the UV residuals
    exist and the UV tests stop on the missing ``expr.data.reg.bio`` (:164).  Here that combination is a ValueError.
  * :287-316  the biological ANOVA sits inside ``if (!is.null(continuous.uv))`` -- always true, since an empty
    selection is ``character(0)``, not NULL -- and uses all samples.  Here it always runs, on all samples.
"""
from __future__ import annotations

import numpy as np

from .api import _printColoredMessage, get_engine

__all__ = ["supervisedFindNCG", "r_rank", "model_matrix"]


def r_rank(x):
    """base::rank(x) with ties.method = "average"."""
    x = np.asarray(x, dtype=np.float64)
    order = np.argsort(x, kind="stable")
    ranks = np.empty(x.shape[0], dtype=np.float64)
    sx = x[order]
    start = 0
    n = x.shape[0]
    while start < n:
        end = start
        while end + 1 < n and sx[end + 1] == sx[start]:
            end += 1
        ranks[order[start:end + 1]] = 0.5 * (start + end) + 1.0
        start = end + 1
    return ranks


def _is_categorical(col):
    return all(isinstance(v, str) for v in col)


def _codes(col, min_count=1):
    """Level codes in sorted level order; levels with fewer than ``min_count`` samples get -1 (left out)."""
    col = np.asarray([str(v) for v in col], dtype=object)
    levels, inv, counts = np.unique(col.astype(str), return_inverse=True, return_counts=True)
    codes = inv.astype(np.int32)
    codes[counts[inv] < min_count] = -1
    return codes


def model_matrix(col_data, variables):
    """model.matrix(~ v1 + v2 + ...) as lm builds it (supervisedFindNCG.R:113-118): an intercept, treatment contrasts (one 0/1
    column per level but the first, levels sorted) for character / factor variables, numeric variables as they are."""
    m = len(next(iter(col_data.values()))) if not variables else len(col_data[variables[0]])
    cols = [np.ones(m)]
    for v in variables:
        col = col_data[v]
        if _is_categorical(col):
            col = np.asarray([str(t) for t in col])
            for lev in np.unique(col)[1:]:
                cols.append((col == lev).astype(np.float64))
        else:
            cols.append(np.asarray(col, dtype=np.float64))
    return np.ascontiguousarray(np.stack(cols, axis=1))


def supervisedFindNCG(se_obj, assay_name, bio_variables, uv_variables, no_ncg=1000, regress_out_uv_variables=False,
                      regress_out_bio_variables=False, apply_normalization=False, normalization="CPM", apply_log=True,
                      pseudo_count=1, corr_method="pearson", a=0.05, rho=0, anova_method="aov", assess_se_obj=True,
                      assess_variables=False, remove_na="both", plot_output=False, fast_pca=True, save_se_obj=True, verbose=True,
                      resident=False):
    """``resident`` (not in R): the assay is the one the engine already holds (uploaded under the current plan, its log
    transform included) -- the statistics run on it in place and nothing crosses PCIe but the per-gene results."""
    _printColoredMessage("------------The supervisedFindNCG function starts:", verbose)
    if apply_normalization:
        raise NotImplementedError("apply.normalization (applyOtherNormalizations, supervisedFindNCG.R:134-146) is outside the CUDA path")
    if regress_out_uv_variables and regress_out_bio_variables:
        raise ValueError("regress.out.uv.variables and regress.out.bio.variables are both TRUE: the reference computes only the "
                         "UV residuals (supervisedFindNCG.R:111-133) and then stops on the missing expr.data.reg.bio (:164)")
    if corr_method not in ("pearson", "spearman"):
        raise NotImplementedError("corr.method = '%s': Pearson and Spearman correlations are built" % corr_method)
    bio_variables = [bio_variables] if isinstance(bio_variables, str) else list(bio_variables or [])
    uv_variables = [uv_variables] if isinstance(uv_variables, str) else list(uv_variables or [])
    eng = get_engine()
    if resident:
        Y = None                                                                  # slot = -1: the resident assay
    else:
        assay = np.asarray(se_obj.assay(assay_name), dtype=np.float64)            # genes x samples
        Y = np.ascontiguousarray(assay.T)
    cat_uv = [v for v in uv_variables if _is_categorical(se_obj.col_data[v])]            # :153-154
    con_uv = [v for v in uv_variables if not _is_categorical(se_obj.col_data[v])]
    cat_bio = [v for v in bio_variables if _is_categorical(se_obj.col_data[v])]          # :246-247
    con_bio = [v for v in bio_variables if not _is_categorical(se_obj.col_data[v])]
    kw = dict(apply_log=apply_log, pseudo_count=pseudo_count)
    all_tests, ranks = {}, []
    plain = dict(X=Y, slot=-1, **kw)
    regressed = dict(X=None, slot=-2)                                            # the residuals ruviii_regress_out left on the device
    if regress_out_uv_variables:                                                 # :111-121: the biology tests run on the residuals
        _printColoredMessage("Regressing out the unwanted variables: %s." % "&".join(uv_variables), verbose)
        eng.regress_out(Y, model_matrix(se_obj.col_data, uv_variables), **kw)
    elif regress_out_bio_variables:                                              # :122-133: the unwanted-variation tests do
        _printColoredMessage("Regressing out the biological variables: %s." % "&".join(bio_variables), verbose)
        eng.regress_out(Y, model_matrix(se_obj.col_data, bio_variables), **kw)
    uv_src = regressed if regress_out_bio_variables else plain                   # :163-167, :199-203
    bio_src = regressed if regress_out_uv_variables else plain                   # :248-254, :288-294

    def corr(v, src):                             # correls(y, x, type = corr.method)[, "correlation"], :214-229 / :270-285
        cov = np.asarray(se_obj.col_data[v], dtype=np.float64)
        if corr_method == "spearman":             # ranks are invariant under log2(x + pseudo.count)
            return eng.gene_spearman(src["X"], covariate=cov, slot=src["slot"])["rho"]
        return eng.gene_stats(covariate=cov, **src)["r"]

    _printColoredMessage("### Finding genes that are highly affected by unwanted variation:", verbose)
    for v in cat_uv:                                                                     # :179-191
        F = eng.gene_stats(groups=_codes(se_obj.col_data[v], min_count=3), **uv_src)["F"]
        all_tests.setdefault("anova.gene.uv", {})[v] = dict(statistic=F, ranked_genes=r_rank(F))
        ranks.append(all_tests["anova.gene.uv"][v]["ranked_genes"])
    for v in con_uv:                                                                     # :214-229
        r = np.round(corr(v, uv_src), 2)
        all_tests.setdefault("corr.genes.uv", {})[v] = dict(correlation=r, ranked_genes=r_rank(np.abs(r)))
        ranks.append(all_tests["corr.genes.uv"][v]["ranked_genes"])
    _printColoredMessage("### Finding genes that are not highly affected by biology", verbose)
    for v in con_bio:                                                                    # :270-285
        r = np.round(corr(v, bio_src), 2)
        all_tests.setdefault("corr.genes.bio", {})[v] = dict(correlation=r, ranked_genes=r_rank(-np.abs(r)))
        ranks.append(all_tests["corr.genes.bio"][v]["ranked_genes"])
    for v in cat_bio:                                                                    # :302-316
        F = eng.gene_stats(groups=_codes(se_obj.col_data[v]), **bio_src)["F"]
        all_tests.setdefault("anova.gene.bio", {})[v] = dict(statistic=F, ranked_genes=r_rank(-F))
        ranks.append(all_tests["anova.gene.bio"][v]["ranked_genes"])
    if not ranks:
        raise ValueError("no bio.variables / uv.variables to test")
    prod = np.prod(np.stack(ranks, axis=1), axis=1)                                      # :342-356
    ncg_selected = r_rank(-prod) < no_ncg + 1                                            # :357
    if save_se_obj:
        se_obj.metadata["NCG"] = ncg_selected                                            # :437
        se_obj.metadata.setdefault("NCG.tests", all_tests)
        _printColoredMessage("The NCG are saved to metadata@NCG.", verbose)
        return se_obj
    return ncg_selected
