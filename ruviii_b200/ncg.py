"""Negative-control-gene selection -- the other step immediately before the RUV-III path.

Host side of ``supervisedFindNCG`` (R/supervisedFindNCG.R:44-461) with ``regress.out.* = FALSE`` and
``apply.normalization = FALSE`` (the reference itself marks the regress.out branch "NEED to BE tEStED", :110), for
``corr.method`` "pearson" (default) and "spearman": the same arguments, the same sequence of tests, ranks and rank
product.  The per-gene statistics -- one-way ANOVA F across every categorical variable, Pearson correlation with every
continuous one -- are the streaming kernel K11 through ``ruviii_gene_stats``; Spearman's rho is K11b
(``ruviii_gene_spearman``: a shared-memory sort of every gene's samples);
``rank()``, the ``round(., 2)`` the reference applies to the correlations, and the rank product are O(genes) host
bookkeeping.  Plotting (:359-425) is not reproduced.

Reference quirks kept or resolved (stated here because they change which genes come out):
  * :183-187  levels of a categorical UV variable with fewer than three samples are dropped from the data, but the
    grouping vector handed to row_oneway_equalvar is not subset (an error in R whenever a level is dropped).  Here
    the samples of such levels are left out of both.
  * :287-316  the biological ANOVA sits inside ``if (!is.null(continuous.uv))`` -- always true, since an empty
    selection is ``character(0)``, not NULL -- and uses all samples.  Here it always runs, on all samples.
"""
from __future__ import annotations

import numpy as np

from .api import _printColoredMessage, get_engine

__all__ = ["supervisedFindNCG", "r_rank"]


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


def supervisedFindNCG(se_obj, assay_name, bio_variables, uv_variables, no_ncg=1000, regress_out_uv_variables=False,
                      regress_out_bio_variables=False, apply_normalization=False, normalization="CPM", apply_log=True,
                      pseudo_count=1, corr_method="pearson", a=0.05, rho=0, anova_method="aov", assess_se_obj=True,
                      assess_variables=False, remove_na="both", plot_output=False, fast_pca=True, save_se_obj=True, verbose=True,
                      resident=False):
    """``resident`` (not in R): the assay is the one the engine already holds (uploaded under the current plan, its log
    transform included) -- the statistics run on it in place and nothing crosses PCIe but the per-gene results."""
    _printColoredMessage("------------The supervisedFindNCG function starts:", verbose)
    if regress_out_uv_variables or regress_out_bio_variables or apply_normalization:
        raise NotImplementedError("regress.out.* / apply.normalization (supervisedFindNCG.R:111-146) are outside the CUDA path")
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

    def corr(v):                                  # correls(y, x, type = corr.method)[, "correlation"], :214-229 / :270-285
        cov = np.asarray(se_obj.col_data[v], dtype=np.float64)
        if corr_method == "spearman":             # ranks are invariant under log2(x + pseudo.count)
            return eng.gene_spearman(Y, covariate=cov)["rho"]
        return eng.gene_stats(Y, covariate=cov, **kw)["r"]

    _printColoredMessage("### Finding genes that are highly affected by unwanted variation:", verbose)
    for v in cat_uv:                                                                     # :179-191
        F = eng.gene_stats(Y, groups=_codes(se_obj.col_data[v], min_count=3), **kw)["F"]
        all_tests.setdefault("anova.gene.uv", {})[v] = dict(statistic=F, ranked_genes=r_rank(F))
        ranks.append(all_tests["anova.gene.uv"][v]["ranked_genes"])
    for v in con_uv:                                                                     # :214-229
        r = np.round(corr(v), 2)
        all_tests.setdefault("corr.genes.uv", {})[v] = dict(correlation=r, ranked_genes=r_rank(np.abs(r)))
        ranks.append(all_tests["corr.genes.uv"][v]["ranked_genes"])
    _printColoredMessage("### Finding genes that are not highly affected by biology", verbose)
    for v in con_bio:                                                                    # :270-285
        r = np.round(corr(v), 2)
        all_tests.setdefault("corr.genes.bio", {})[v] = dict(correlation=r, ranked_genes=r_rank(-np.abs(r)))
        ranks.append(all_tests["corr.genes.bio"][v]["ranked_genes"])
    for v in cat_bio:                                                                    # :302-316
        F = eng.gene_stats(Y, groups=_codes(se_obj.col_data[v]), **kw)["F"]
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
