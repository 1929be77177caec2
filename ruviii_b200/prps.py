"""PRPS construction -- the step immediately before the RUV-III path.

Host side of ``prpsForCategoricalUV`` (R/prpsForCategoricalUV.R:36-215), ``prpsForContinuousUV``
(R/prpsForContinuousUV.R:47-262) and their driver ``supervisedPRPS`` (R/supervisedPRPS.R:112-211): the same
arguments, error messages and PRPS set names.  What stays on the host is the bookkeeping on ``colData`` (which
samples share a biology x batch cell; which sit at the top / bottom of a continuous variable); every pseudo-sample
``rowMeans(expre.data[, index.sample])`` is computed by the segmented-mean kernel K0 through
``ruviii_prps`` (include/ruviii.h).  Nothing here averages on the CPU.

Orientation: R keeps a PRPS block as genes x pseudo-samples with the set ids as column names; here it is a
:class:`ReplicateData` (pseudo-samples x genes, row-major) -- the same bytes, and already ``t(...)`` as
``normalise.R:143`` wants it.
"""
from __future__ import annotations

import numpy as np

from .api import ReplicateData, _printColoredMessage, get_engine

__all__ = ["prps_sets_categorical", "prps_sets_continuous", "prpsForCategoricalUV", "prpsForContinuousUV",
           "supervisedPRPS", "replicate_data_from_metadata"]


def _labels(values):
    """gsub('_', '-', x) on a character / factor colData column (prpsForCategoricalUV.R:82-85)."""
    return np.asarray([str(v).replace("_", "-") for v in values], dtype=object)


def prps_sets_categorical(bio, uv, uv_variable, min_sample_for_prps=3):
    """Index sets of prpsForCategoricalUV.R:99-139.

    ``table(bio, uv)`` (levels in sorted order), keep the biological groups that reach ``min.sample.for.prps``
    samples in more than one batch (:103-104), and inside each of them make one pseudo-sample per qualifying
    batch (:123-132); all pseudo-samples of a biological group share the set id ``<bio>||<uv.variable>`` (:134-136).
    Returns (list of int32 index arrays, list of set ids)."""
    bio, uv = _labels(bio), _labels(uv)
    bio_levels, uv_levels = sorted(set(bio)), sorted(set(uv))
    table = np.array([[int(np.sum((bio == b) & (uv == u))) for u in uv_levels] for b in bio_levels])
    keep = (table >= min_sample_for_prps).sum(axis=1) > 1
    if not keep.any():
        raise ValueError("There are not enough samples to create PRPS across the batches.")       # :106-108
    sets, names = [], []
    for bi in np.nonzero(keep)[0]:
        for ui in np.nonzero(table[bi] >= min_sample_for_prps)[0]:
            sets.append(np.nonzero((bio == bio_levels[bi]) & (uv == uv_levels[ui]))[0].astype(np.int32))
            names.append("%s||%s" % (bio_levels[bi], uv_variable))
    return sets, names


def prps_sets_continuous(uv_values, bio, batch, min_sample_for_prps=3, min_sample_per_batch=10):
    """Index sets of prpsForContinuousUV.R:165-201.

    Samples are grouped by ``paste0(batch, "_", bio)``; groups with fewer than ``min.sample.per.batch`` samples are
    dropped (:172-175); in every remaining group the ``min.sample.for.prps`` samples with the smallest and with the
    largest value of the continuous variable form the two pseudo-samples of the set ``<batch>_<bio>-LS``
    (:176-201; dplyr's arrange is stable, groups come out in sorted key order).  Returns (sets, set ids), the
    low pseudo-sample before the high one as ``cbind(rowMeans(bot), rowMeans(top))`` orders them."""
    uv_values = np.asarray(uv_values, dtype=np.float64)
    key = np.asarray(["%s_%s" % (b, g) for b, g in zip(batch, bio)], dtype=object)
    levels, counts = np.unique(key.astype(str), return_counts=True)
    sets, names = [], []
    asc = np.argsort(uv_values, kind="stable")
    desc = np.argsort(-uv_values, kind="stable")
    for lev, cnt in zip(levels, counts):
        if cnt < min_sample_per_batch:
            continue
        bot = [i for i in asc if key[i] == lev][:min_sample_for_prps]
        top = [i for i in desc if key[i] == lev][:min_sample_for_prps]
        sets += [np.asarray(bot, dtype=np.int32), np.asarray(top, dtype=np.int32)]
        names += ["%s-LS" % lev] * 2
    if not sets:
        raise ValueError("There is not enough samples to create PRPS for contenious sources of unwanted variation.")
    return sets, names


def _store(se_obj, bio_variable, uv_variable, assay_name, block):
    key = "bio:%s||uv:%s||data:%s" % (bio_variable, uv_variable, assay_name)                 # prpsForCategoricalUV.R:176-189
    se_obj.metadata.setdefault("PRPS", {}).setdefault("supervised", {})[key] = block
    return key


def prpsForCategoricalUV(se_obj, assay_name, uv_variable, bio_variable, min_sample_for_prps=3, apply_log=True,
                         pseudo_count=1, remove_na="both", assess_se_obj=True, save_se_obj=True, verbose=True, sets_only=False):
    """R/prpsForCategoricalUV.R:36-215.  ``sets_only`` (not in R): return ``(index sets, set ids)`` -- the colData bookkeeping
    alone -- for callers that let the engine build the pseudo-samples from an assay it already holds (normalise())."""
    _printColoredMessage("------------The prpsForCategoricalUV function starts:", verbose)
    if isinstance(assay_name, (list, tuple)):
        raise ValueError("The function can only take a single assay.name.")
    if isinstance(uv_variable, (list, tuple)):
        raise ValueError("The function can only take a single categorical variable for the uv.variable argument.")
    if uv_variable is None:
        raise ValueError("The uv.variable cannot be empty.")
    uv = se_obj.col_data[uv_variable]
    if not all(isinstance(v, str) for v in uv):
        raise ValueError("The uv.variable should be a categorical source of unwanted variation, e.g., time effects")
    if len(set(uv)) == 1:
        raise ValueError("The uv.variable should have at least two levels.")
    bio = se_obj.col_data[bio_variable]
    if not all(isinstance(v, str) for v in bio):
        raise ValueError("The bio.variable should be a categorical biological variable, e.g., cancer subtypes")
    if len(set(bio)) == 1:
        raise ValueError("The bio.variable should have at least two groups.")
    sets, names = prps_sets_categorical(bio, uv, uv_variable, min_sample_for_prps)
    if sets_only:
        return sets, names
    assay = np.asarray(se_obj.assay(assay_name), dtype=np.float64)
    values = get_engine().prps(assay.T, sets, apply_log=apply_log, pseudo_count=pseudo_count)   # K0 on the device
    block = ReplicateData(values, names)
    _printColoredMessage("%d sets of PRPS with the total number of %d pseudo-samples are created between different "
                         "batches of the %s variable." % (len(set(names)), len(names), uv_variable), verbose)
    if save_se_obj:
        _store(se_obj, bio_variable, uv_variable, assay_name, block)
        return se_obj
    return block


def prpsForContinuousUV(se_obj, assay_name, apply_log=True, pseudo_count=1, uv_variable=None, bio_variable=None,
                        batch_variable=None, min_sample_for_prps=3, min_sample_per_batch=10, assess_se_obj=True,
                        remove_na="both", save_se_obj=True, verbose=True, sets_only=False):
    """R/prpsForContinuousUV.R:47-262.  ``sets_only``: see :func:`prpsForCategoricalUV`."""
    _printColoredMessage("------------The prpsForContinuousUV function starts.", verbose)
    if isinstance(assay_name, (list, tuple)):
        raise ValueError("The function can only take a single assay.name.")
    if uv_variable is None:
        raise ValueError("The uv.variable cannot be empty.")
    if batch_variable is None:
        # sample.annot[, c('sOrder', uv, bio, NULL)] has three columns and then receives four names (:169-170)
        raise ValueError("'names' attribute [4] must be the same length as the vector [3] (batch.variable = NULL)")
    sets, names = prps_sets_continuous(se_obj.col_data[uv_variable], se_obj.col_data[bio_variable],
                                       se_obj.col_data[batch_variable], min_sample_for_prps, min_sample_per_batch)
    if sets_only:
        return sets, names
    assay = np.asarray(se_obj.assay(assay_name), dtype=np.float64)
    values = get_engine().prps(assay.T, sets, apply_log=apply_log, pseudo_count=pseudo_count)
    block = ReplicateData(values, names)
    _printColoredMessage("In total %d PRPS sets with %d the total number of pseudo-samples are created for the removal of "
                         "%s effects." % (len(set(names)), len(names), uv_variable), verbose)
    if save_se_obj:
        _store(se_obj, bio_variable, uv_variable, assay_name, block)
        return se_obj
    return block


def supervisedPRPS(se_obj, assay_name, bio_variable, uv_variables, batch_variable=None, min_sample_for_prps=3,
                   min_sample_per_batch=6, apply_log=True, pseudo_count=1, assess_se_obj=True,
                   assess_cor_variables=False, remove_na="both", save_se_obj=True, verbose=True, sets_only=False):
    """R/supervisedPRPS.R:62-260: one PRPS block per unwanted-variation variable, categorical ones (character columns)
    through ``prpsForCategoricalUV``, numeric ones through ``prpsForContinuousUV``.  ``sets_only``: dict block key ->
    (index sets, set ids), nothing computed."""
    if uv_variables is None:
        raise ValueError("The uv.variables cannot be empty.")
    uv_variables = [uv_variables] if isinstance(uv_variables, str) else list(uv_variables)
    blocks = {}
    for uv in uv_variables:
        col = se_obj.col_data[uv]
        if all(isinstance(v, str) for v in col):
            out = prpsForCategoricalUV(se_obj, assay_name, uv, bio_variable, min_sample_for_prps, apply_log, pseudo_count,
                                       remove_na, False, save_se_obj, verbose, sets_only)
        else:
            out = prpsForContinuousUV(se_obj, assay_name, apply_log, pseudo_count, uv, bio_variable, batch_variable,
                                      min_sample_for_prps, min_sample_per_batch, False, remove_na, save_se_obj, verbose, sets_only)
        if sets_only:
            blocks["bio:%s||uv:%s||data:%s" % (bio_variable, uv, assay_name)] = out
            continue
        if not save_se_obj:
            blocks["bio:%s||uv:%s||data:%s" % (bio_variable, uv, assay_name)] = out
    return blocks if sets_only else (se_obj if save_se_obj else blocks)


def replicate_data_from_metadata(se_obj):
    """``t(do.call(cbind, se.obj@metadata[['PRPS']][['supervised']]))`` (normalise.R:143)."""
    blocks = list(se_obj.metadata["PRPS"]["supervised"].values())
    return ReplicateData(np.vstack([b.values for b in blocks]), [nm for b in blocks for nm in b.row_names])
