"""computePCA -- the step right after the RUV-III path (normAssessment.R:113 runs it on every ``RUV_K*``
assay).  Host side of ``computePCA`` (R/computePCA.R:38-262): same arguments, messages and result layout
(``metadata$metric[[assay]]$fastPCA`` or ``$PCA`` ``= list(sing.val = list(u, d, v), variation)``); the decomposition is
``ruviii_pca`` (centre -> Gram -> eigensolver -> projection on the device).  ``fast.pca = TRUE`` asks for the ``nb.pcs``
leading triplets (:120-134); ``fast.pca = FALSE`` (:150-190, ``svd(scale(t(x)))``) for all min(samples, genes) of them.
The centred matrix has rank <= samples - 1, so its last singular value is rounding noise and the vectors that go with it are
as arbitrary here as they are in LAPACK.
"""
from __future__ import annotations

import numpy as np

from .api import _printColoredMessage, get_engine

__all__ = ["computePCA"]


def _round_half_even_1(x):
    # R's round(x, 1) (IEC 60559 "round half to even" on the decimal representation) == numpy's rounding rule
    return np.round(x, 1)


def computePCA(se_obj, assay_names="All", apply_log=True, fast_pca=True, nb_pcs=10, return_pc_percentage=True, scale=False,
               center=True, BSPARAM=None, save_se_obj=True, assess_se_obj=True, verbose=True, pseudo_count=1):
    _printColoredMessage("------------The computePCA function starts:", verbose)
    if fast_pca and (nb_pcs is None or nb_pcs == 0):                                          # computePCA.R:55-58
        raise ValueError("To perform fast PCA, the number of PCs (nb.pcs) must specified.")
    if assay_names is None:
        raise ValueError("Please provide at least an assay name.")
    if isinstance(assay_names, str):
        assay_names = list(se_obj.assays) if assay_names == "All" else [assay_names]
    assay_names = sorted(set(assay_names))                                                     # levels(as.factor(...)), :93-97
    eng = get_engine()
    results = {}
    for name in assay_names:
        assay = np.asarray(se_obj.assay(name), dtype=np.float64)                               # genes x samples
        k = int(nb_pcs) if fast_pca else min(assay.shape)                                      # :150-190: svd() keeps every triplet
        res = eng.pca(assay.T, k=k, apply_log=apply_log, pseudo_count=pseudo_count, center=center, scale=scale)
        sv = dict(u=res["u"], d=res["d"], v=res["v"])                                          # rownames: samples / genes
        out = dict(sing_val=sv)
        if return_pc_percentage:                                                               # :137-141
            pct = sv["d"] ** 2 / np.sum(sv["d"] ** 2) * 100.0
            out["variation"] = _round_half_even_1(pct)
        results[name] = out
    if not save_se_obj:
        return results
    for name in assay_names:                                                                   # :194-236
        key = "fastPCA" if fast_pca else "PCA"                                                 # :214-236
        se_obj.metadata.setdefault("metric", {}).setdefault(name, {})[key] = results[name]
        _printColoredMessage("The PCA results are saved to metadata@metric$%s$%s." % (name, key), verbose)
    _printColoredMessage("------------The computePCA function finished.", verbose)
    return se_obj
