"""Host-side mirror of the reference's R interface for the RUV-III hot path.

Same function names, argument meaning, defaults, return shapes and error behaviour as
``ruvIII`` (R/ruvIII.R:38-203), ``fastruvIII`` (R/fastruvIII.R:40-234) and ``ruvIIIMultipleK``
(R/ruvIIIMultipleK.R:34-187); only the arithmetic of ruvIII.R:116-149 is replaced by one call into
libruviii_b200.so.  R's dots become underscores (``assay.name`` -> ``assay_name``).  Because there is no
SummarizedExperiment class in Python, :class:`SummarizedExperiment` below carries just what the path touches:
named assays (genes x samples), sample and gene names, and the ``metadata`` list.

Nothing here computes on the CPU: if the CUDA library or an sm_100 GPU is missing the calls raise.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field

import numpy as np

from . import _cabi
from ._cabi import MODE_FAST, MODE_STACKED, Engine, RuviiiError, make_params

__all__ = ["SummarizedExperiment", "ReplicateData", "ruvIII", "fastruvIII", "ruvIIIMultipleK", "normalise", "get_engine",
           "set_engine", "group_codes"]


@dataclass
class SummarizedExperiment:
    """The slice of a Bioconductor SummarizedExperiment the hot path reads and writes."""
    assays: dict                     # name -> genes x samples float64 array (Fortran order avoids one copy)
    col_names: list                  # sample ids  (colnames(se.obj))
    row_names: list | None = None    # gene ids
    metadata: dict = field(default_factory=dict)
    col_data: dict = field(default_factory=dict)   # colData(se.obj): variable name -> one value per sample

    def assay(self, name):
        if name not in self.assays:
            raise KeyError("assay '%s' is not in the SummarizedExperiment" % name)   # checkSeObj would stop()
        return self.assays[name]


@dataclass
class ReplicateData:
    """``replicate.data``: pseudo-samples x genes with one PRPS set id per row (normalise.R:143)."""
    values: np.ndarray
    row_names: list


_engine = None


def get_engine():
    """The process-wide engine handle (what the R shim keeps in an external pointer)."""
    global _engine
    if _engine is None:
        _engine = Engine(devices=[0])
    return _engine


def set_engine(engine):
    global _engine
    _engine = engine


def group_codes(names):
    """as.integer(factor(names)) - 1: equal names <=> equal code, codes in sorted level order -- the column order of
    replicate.matrix() (ruvIII.R:101), which is also the row order of the apply.average.rep result (ruvIII.R:149)."""
    _, inv = np.unique(np.asarray(list(names), dtype=object).astype(str), return_inverse=True)
    return inv.astype(np.int32)


def _tological(ctl, n):
    # ruvIII.R:80-84; an index vector is 0-based here
    ctl = np.asarray(ctl)
    if ctl.dtype == np.bool_:
        if ctl.shape[0] != n:
            raise ValueError("logical ctl must have length n")
        return ctl
    out = np.zeros(n, dtype=bool)
    out[ctl.astype(np.int64)] = True
    return out


def _printColoredMessage(message, verbose):
    if verbose:
        print(message)


def _prepare(se_obj, assay_name, replicate_data, ctl, inputcheck, n_rows_for_check):
    if not isinstance(replicate_data, ReplicateData):
        raise TypeError("replicate_data must be a ReplicateData(values, row_names)")
    assay = np.asarray(se_obj.assay(assay_name), dtype=np.float64)
    n, m1 = assay.shape
    PS = np.asarray(replicate_data.values, dtype=np.float64)
    if PS.ndim != 2 or PS.shape[1] != n:
        raise ValueError("replicate.data must have one column per gene of the assay")
    if len(replicate_data.row_names) != PS.shape[0]:
        raise ValueError("replicate.data needs one row name per row")
    ctl = _tological(ctl, n)
    Y = assay.T                                                  # m1 x n row-major view when assay is F-ordered
    if inputcheck:                                               # ruvIII.R:104-114 (warnings only)
        if n_rows_for_check > n:
            warnings.warn("Please ensure that rows correspond to observations (e.g. RNA-seq assay) and columns "
                          "correspond to features (e.g. genes).")
        if np.isnan(Y).any() or np.isnan(PS).any():
            warnings.warn("Y contains missing values. This is not supported.")
        if np.isinf(Y).any() or np.isinf(PS).any():
            warnings.warn("Y contains infinity values. This is not supported.")
    return assay, Y, PS, ctl, n, m1


def _design_matrix_t(eta, n, include_intercept):
    """t(design.matrix(eta, include.intercept)) of ruv::RUV1: q x n with one column per gene.  A single number means
    "intercept only" (RUV1: ``if (length(eta) == 1) eta = matrix(1, n, 1)``); a gene-wise matrix may be given as n x q
    (R's orientation) or q x n; the row of ones is prepended unless the constant vector already lies in the span of the
    design (design.matrix: ``sum(residop(matrix(1, n, 1), A)^2) > 1e-8``).  Host-side bookkeeping on a q x n matrix, q <= 32."""
    e = np.asarray(eta, dtype=np.float64)
    if e.ndim == 0 or e.size == 1:
        return np.ones((1, n))
    e = np.atleast_2d(e)
    if e.shape[1] != n and e.shape[0] == n:
        e = e.T
    if e.shape[1] != n:
        raise ValueError("eta must have one row per gene (n x q)")
    if include_intercept:
        ones = np.ones(n)
        coef = np.linalg.lstsq(e.T, ones, rcond=None)[0]
        if float(np.sum((ones - e.T @ coef) ** 2)) > 1e-8:
            e = np.vstack([np.ones((1, n)), e])
    return np.ascontiguousarray(e)


def _core(mode, se_obj, assay_name, apply_log, pseudo_count, replicate_data, ctl, ks, eta, apply_average_rep,
          fullalpha, return_info, inputcheck, want_ps, include_intercept=True):
    n_ps = len(replicate_data.row_names) if isinstance(replicate_data, ReplicateData) else 0
    m1_guess = np.asarray(se_obj.assay(assay_name)).shape[1]
    assay, Y, PS, ctl, n, m1 = _prepare(se_obj, assay_name, replicate_data, ctl, inputcheck,
                                        n_ps if mode == MODE_FAST else m1_guess + n_ps)
    if mode == MODE_FAST:
        groups = group_codes(list(replicate_data.row_names))                       # fastruvIII.R:117-118
    else:
        groups = group_codes(list(se_obj.col_names) + list(replicate_data.row_names))   # ruvIII.R:93,101
    params = make_params(mode=mode, apply_log=apply_log, pseudo_count=pseudo_count, average_rep=apply_average_rep,
                         n_alpha_rows=-1 if (return_info and mode == MODE_STACKED) else 0,   # fastruvIII.R:171: k rows
                         want_ps_rows=want_ps)
    eta_t = None if eta is None else _design_matrix_t(eta, n, include_intercept)
    eng = get_engine()
    eng.set_eta(eta_t)                                                                  # ruvIII.R:116 RUV1(Y, eta, ctl)
    try:
        return _core_run(eng, params, Y, PS, groups, ctl, ks, m1, n, fullalpha, return_info, want_ps, apply_average_rep)
    finally:
        if eta is not None:
            eng.set_eta(None)


def _core_run(eng, params, Y, PS, groups, ctl, ks, m1, n, fullalpha, return_info, want_ps, apply_average_rep):
    if fullalpha is None:
        res = eng.normalise(params, Y, PS, groups, ctl, ks, want_fullalpha=return_info, want_W=return_info,
                            want_ps=want_ps)
    else:                                                                           # ruvIII.R:138
        eng.plan(params, m1, n, PS.shape[0], groups, ctl, ks)
        eng.upload(Y, PS)
        eng.set_fullalpha(np.asarray(fullalpha, dtype=np.float64))
        eng.run()
        res = eng.download(want_fullalpha=False, want_W=return_info, want_ps=want_ps)
        res["fullalpha"] = np.asarray(fullalpha, dtype=np.float64)
        if apply_average_rep:
            res["averaged"] = eng.download_averaged()
    return res, m1, n


def _save(se_obj, assay_name, k, newY_block, return_info, W, fullalpha):
    new_assay_name = "RUV_K%d_on_%s" % (k, assay_name)               # ruvIII.R:154
    se_obj.assays[new_assay_name] = newY_block.T                      # :158  t(newY[1:m1, ])  (genes x samples)
    if return_info:                                                   # :165-189
        se_obj.metadata.setdefault("RUVIII", {}).setdefault(new_assay_name, {})
        se_obj.metadata["RUVIII"][new_assay_name]["W"] = W
        se_obj.metadata["RUVIII"][new_assay_name]["fullalpha"] = fullalpha
    return new_assay_name


def _single(mode, se_obj, assay_name, apply_log, pseudo_count, replicate_data, ctl, k, eta, include_intercept,
            apply_average_rep, fullalpha, return_info, inputcheck, save_se_obj, verbose):
    name = "ruvIII"
    _printColoredMessage("------------The %s function starts:" % name, verbose)
    if k is None:
        raise ValueError("k cannot be 0. This means no adjustment will be made.")   # ruvIII.R:62-63
    if mode == MODE_FAST:
        _, counts = np.unique(np.asarray(replicate_data.row_names), return_counts=True)
        if counts.min() == 1:                                                        # fastruvIII.R:67-68
            raise ValueError("There are only replicated samples of a single sample in the replicate.data")
    want_ps = (mode == MODE_STACKED) and not save_se_obj     # t(newY) / newY carry the PRPS rows too (ruvIII.R:192,195)
    _printColoredMessage("### Performing the RUVIII normalization", verbose)
    res, m1, n = _core(mode, se_obj, assay_name, apply_log, pseudo_count, replicate_data, ctl, [int(k)], eta,
                       apply_average_rep, fullalpha, return_info, inputcheck, want_ps, include_intercept)
    info = res["info"]
    undefined = info.no_replicates                   # ruvIII.R:128-129 leaves W / fullalpha undefined
    W = None if res["W"] is None else res["W"][0]
    fa = res["fullalpha"]
    if return_info and undefined:
        raise RuntimeError("object 'W' not found: ncol(M) >= m, no replicates (ruvIII.R:128-129 with return.info)")
    if save_se_obj:
        block = res["averaged"][0][:m1] if apply_average_rep else res["newY"][0]     # t(newY[1:ncol(expr.data), ])
        new_name = _save(se_obj, assay_name, int(k), block, return_info, W, fa)
        _printColoredMessage("The normalized data %s is saved into a new assay of the SummarizedExperiment object."
                             % new_name, verbose)
        return se_obj
    newY = res["newY"][0] if res["newY_ps"] is None else np.vstack([res["newY"][0], res["newY_ps"][0]])
    if apply_average_rep:
        newY = res["averaged"][0]                                                    # ruvIII.R:148-149: p rows
    if not return_info:
        return newY.T                                                                # ruvIII.R:191-192  t(newY)
    return dict(newY=newY, fullalpha=fa, W=W)                                        # ruvIII.R:193-197


def ruvIII(se_obj, assay_name, apply_log=True, pseudo_count=1, replicate_data=None, ctl=None, k=None, eta=None,
           include_intercept=True, apply_average_rep=False, fullalpha=None, return_info=False, inputcheck=True,
           assess_se_obj=True, remove_na="measurements", save_se_obj=True, verbose=True):
    """RUV-III with PRPS (R/ruvIII.R:38-203).  Returns the SummarizedExperiment with the new assay
    ``RUV_K<k>_on_<assay>`` (save_se_obj), ``t(newY)`` (genes x (samples + PRPS), :191-192), or
    ``dict(newY, fullalpha, W)`` (:193-197)."""
    return _single(MODE_STACKED, se_obj, assay_name, apply_log, pseudo_count, replicate_data, ctl, k, eta,
                   include_intercept, apply_average_rep, fullalpha, return_info, inputcheck, save_se_obj, verbose)


def fastruvIII(se_obj, assay_name, apply_log=True, pseudo_count=1, replicate_data=None, ctl=None, BSPARAM=None,
               k=None, eta=None, include_intercept=True, apply_average_rep=False, fullalpha=None, return_info=False,
               inputcheck=True, assess_se_obj=True, remove_na="measurements", save_se_obj=True, verbose=True):
    """Fast RUV-III (R/fastruvIII.R:40-234): SVD on the PRPS rows only, mu over the samples only.
    ``BSPARAM`` is accepted for signature compatibility; the engine always computes the exact decomposition
    (the default ``bsparam()`` = ExactParam)."""
    return _single(MODE_FAST, se_obj, assay_name, apply_log, pseudo_count, replicate_data, ctl, k, eta,
                   include_intercept, apply_average_rep, fullalpha, return_info, inputcheck, save_se_obj, verbose)


def ruvIIIMultipleK(se_obj, assay_name, apply_log=True, pseudo_count=1, replicate_data=None, ctl=None, k=None,
                    eta=None, include_intercept=True, apply_average_rep=False, fullalpha=None, return_info=False,
                    inputcheck=True, assess_se_obj=True, remove_na="measurements", save_se_obj=True, verbose=True):
    """RUV-III for a vector of k (R/ruvIIIMultipleK.R:34-187).  The reference runs a complete ``ruvIII`` per k
    (:75-94); here one residop / Gram / eigensolve serves every k and the update kernel emits all of them in one
    pass over Y.  Return shapes follow the three branches (:74-98, :101-154, :157-185); the two list-returning
    branches are given the ``ctl`` the reference forgets to forward."""
    _printColoredMessage("------------The ruvIIIMultipleK function starts:", verbose)
    if k is None:
        raise ValueError("k cannot be 0. This means no adjustment will be made.")   # ruvIIIMultipleK.R:57-58
    ks = [int(x) for x in np.atleast_1d(k)]
    want_ps = not save_se_obj
    res, m1, n = _core(MODE_STACKED, se_obj, assay_name, apply_log, pseudo_count, replicate_data, ctl, ks, eta,
                       apply_average_rep, fullalpha, return_info, inputcheck, want_ps, include_intercept)
    if return_info and res["info"].no_replicates:
        raise RuntimeError("object 'W' not found: ncol(M) >= m, no replicates (ruvIII.R:128-129 with return.info)")
    Ws = res["W"] if res["W"] is not None else [None] * len(ks)
    if save_se_obj:                                                                  # :74-98
        for i, kv in enumerate(ks):
            block = res["averaged"][i][:m1] if apply_average_rep else res["newY"][i]
            _save(se_obj, assay_name, kv, block, return_info, Ws[i], res["fullalpha"])
        _printColoredMessage("------------The ruvIIIMultipleK function finished:", verbose)
        return se_obj
    if not return_info:                                                              # :101-154
        out = {}
        for i, kv in enumerate(ks):
            label = ("RUVIII_K:%d_Data:%s" if i == 0 else "RUV_K%d_Data:%s") % (kv, assay_name)   # :121,146
            out[label] = (res["averaged"][i] if apply_average_rep else np.vstack([res["newY"][i], res["newY_ps"][i]])).T
        return out
    out = []                                                                         # :157-185
    for i, kv in enumerate(ks):
        newY = res["newY"][i] if res["newY_ps"] is None else np.vstack([res["newY"][i], res["newY_ps"][i]])
        if apply_average_rep:
            newY = res["averaged"][i]
        out.append(dict(newY=newY, fullalpha=res["fullalpha"], W=Ws[i]))
    return out


def normalise(se_obj, assay_name, apply_log=True, pseudo_count=1, bio_variable_prps=None, uv_variables_prps=None,
              batch_variable_prps=None, min_sample_for_prps=3, min_sample_per_batch_prps=6, assess_cor_variables_prps=False,
              norm_assay_name_ncg=None, apply_normalization_ncg=False, normalization_ncg="CPM", bio_variables_ncg=None,
              uv_variables_ncg=None, no_ncg=1000, regress_out_uv_variables_ncg=False, regress_out_bio_variables_ncg=False,
              k=None, eta=None, include_intercept=True, apply_average_rep=False, fullalpha=None, inputcheck=True,
              assess_se_obj=True, remove_na="both", verbose=True):
    """R/normalise.R:61-167: supervisedPRPS -> supervisedFindNCG -> ruvIIIMultipleK.

    PRPS construction (:mod:`ruviii_b200.prps`), the negative-control gene statistics (:mod:`ruviii_b200.ncg`) and the
    RUV-III step all run on the device.  ``norm_assay_name_ncg = None`` with ``metadata$NCG`` already present skips the
    selection (an extension: the reference stops when the argument is missing, normalise.R:103-106)."""
    from . import prps
    _printColoredMessage("------------The normalise function starts.", verbose)
    if k is None:
        raise ValueError("k cannot be 0. This means no adjustment will be made.")                     # normalise.R:95-96
    if assay_name is None:
        raise ValueError("No assay name has been provided.")
    if isinstance(assay_name, (list, tuple)):
        raise ValueError("The function can only take a single assay.name.")
    # One pipeline, one upload (normalise.R:111,127,143-161): when the three steps work on the same assay the engine keeps it on
    # the device -- the pseudo-samples are built there from the index sets (K0), the control-gene statistics run on it in place
    # (K11), the plan is redone with the selected ctl (keep_resident) and RUV-III follows; the assay crosses PCIe once.
    import os
    if (eta is None and fullalpha is None and not apply_average_rep and norm_assay_name_ncg in (None, assay_name)
            and not (regress_out_uv_variables_ncg or regress_out_bio_variables_ncg or apply_normalization_ncg)
            and os.environ.get("RUVIII_NORMALISE_RESIDENT", "1") != "0"):
        return _normalise_resident(se_obj, assay_name, apply_log, pseudo_count, bio_variable_prps, uv_variables_prps,
                                   batch_variable_prps, min_sample_for_prps, min_sample_per_batch_prps, norm_assay_name_ncg,
                                   bio_variables_ncg, uv_variables_ncg, no_ncg, k, inputcheck, assess_se_obj, verbose)
    se_obj = prps.supervisedPRPS(se_obj, assay_name, bio_variable_prps, uv_variables_prps, batch_variable_prps,
                                 min_sample_for_prps, min_sample_per_batch_prps, apply_log, pseudo_count, assess_se_obj,
                                 assess_cor_variables_prps, save_se_obj=True, verbose=verbose)          # :110-124
    if norm_assay_name_ncg is None and "NCG" not in se_obj.metadata:
        raise ValueError("norm.assay.name.ncg is empty. Please select an assay to use for the Negative Control Genes selection")
    if norm_assay_name_ncg is not None:                                                               # :127-141
        from .ncg import supervisedFindNCG
        se_obj = supervisedFindNCG(se_obj, norm_assay_name_ncg, bio_variables_ncg, uv_variables_ncg, no_ncg=no_ncg,
                                   regress_out_uv_variables=regress_out_uv_variables_ncg,
                                   regress_out_bio_variables=regress_out_bio_variables_ncg,
                                   apply_normalization=apply_normalization_ncg, normalization=normalization_ncg,
                                   apply_log=apply_log, pseudo_count=pseudo_count, assess_se_obj=assess_se_obj,
                                   save_se_obj=True, verbose=verbose)
    replicate_data = prps.replicate_data_from_metadata(se_obj)                                        # :143
    se_obj = ruvIIIMultipleK(se_obj, assay_name, apply_log=apply_log, pseudo_count=pseudo_count,
                             replicate_data=replicate_data, ctl=se_obj.metadata["NCG"], k=k, eta=eta,
                             include_intercept=include_intercept, apply_average_rep=apply_average_rep, fullalpha=fullalpha,
                             return_info=False, inputcheck=inputcheck, assess_se_obj=assess_se_obj,
                             remove_na="measurements", save_se_obj=True, verbose=verbose)                 # :144-161
    _printColoredMessage("------------The normalise function finished.", verbose)
    return se_obj


def _normalise_resident(se_obj, assay_name, apply_log, pseudo_count, bio_variable_prps, uv_variables_prps, batch_variable_prps,
                        min_sample_for_prps, min_sample_per_batch_prps, norm_assay_name_ncg, bio_variables_ncg, uv_variables_ncg,
                        no_ncg, k, inputcheck, assess_se_obj, verbose):
    """normalise() with the assay resident on the device from the first step to the last (see :func:`normalise`)."""
    from . import prps
    from .ncg import supervisedFindNCG
    blocks = prps.supervisedPRPS(se_obj, assay_name, bio_variable_prps, uv_variables_prps, batch_variable_prps, min_sample_for_prps,
                                 min_sample_per_batch_prps, apply_log, pseudo_count, assess_se_obj, False, save_se_obj=True,
                                 verbose=verbose, sets_only=True)                                       # normalise.R:110-124
    sets = [ix for b in blocks.values() for ix in b[0]]
    set_names = [nm for b in blocks.values() for nm in b[1]]
    assay = np.asarray(se_obj.assay(assay_name), dtype=np.float64)                                      # genes x samples
    n, m1 = assay.shape
    ks = [int(x) for x in np.atleast_1d(k)]
    groups = group_codes(list(se_obj.col_names) + set_names)                                            # ruvIII.R:93,101
    eng = get_engine()
    eng.set_prps_sets(sets)
    try:
        eng.plan(make_params(mode=MODE_STACKED, apply_log=apply_log, pseudo_count=pseudo_count), m1, n, len(sets), groups,
                 np.zeros(n, dtype=bool), ks)                                # provisional plan: no control genes known yet
        eng.upload(assay.T, None)                                            # the ONE upload; PS = NULL: built on the device
        if norm_assay_name_ncg is None and "NCG" not in se_obj.metadata:
            raise ValueError("norm.assay.name.ncg is empty. Please select an assay to use for the Negative Control Genes selection")
        if norm_assay_name_ncg is not None:                                                             # :127-141
            se_obj = supervisedFindNCG(se_obj, norm_assay_name_ncg, bio_variables_ncg, uv_variables_ncg, no_ncg=no_ncg,
                                       apply_log=apply_log, pseudo_count=pseudo_count, assess_se_obj=assess_se_obj,
                                       save_se_obj=True, verbose=verbose, resident=True)
        ctl = _tological(se_obj.metadata["NCG"], n)                                                     # ruvIII.R:80-84,102
        eng.plan(make_params(mode=MODE_STACKED, apply_log=apply_log, pseudo_count=pseudo_count, keep_resident=True), m1, n,
                 len(sets), groups, ctl, ks)
        eng.run()
        res = eng.download()
        PS = eng.download_replicate_data()
    finally:
        eng.set_prps_sets([])
    # metadata$PRPS$supervised[[key]]: one block per unwanted-variation variable (prpsForCategoricalUV.R:176-189)
    row = 0
    for key, (bsets, bnames) in blocks.items():
        se_obj.metadata.setdefault("PRPS", {}).setdefault("supervised", {})[key] = ReplicateData(PS[row:row + len(bsets)], bnames)
        row += len(bsets)
    for i, kv in enumerate(ks):                                                                         # ruvIIIMultipleK.R:75-94
        _save(se_obj, assay_name, kv, res["newY"][i], False, None, None)
    _printColoredMessage("------------The normalise function finished.", verbose)
    return se_obj
