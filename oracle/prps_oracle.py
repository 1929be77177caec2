"""CPU ORACLE (test infrastructure, never shipped, never imported by ruviii_b200/) for PRPS construction:
a plain numpy restatement of

  * ``prpsForCategoricalUV``  R/prpsForCategoricalUV.R:82-139
  * ``prpsForContinuousUV``   R/prpsForContinuousUV.R:138-201

PARITY UNPINNED against R: the reference ships no tests or golden vectors for these functions and R is not installed;
``table``/``dplyr`` semantics (sorted levels, stable ``arrange``, groups in sorted key order) are restated from their
documentation.  Returns genes x pseudo-samples matrices with the set ids as column names, as the R functions do.
"""
import numpy as np


def _expre(assay, apply_log, pseudo_count):
    assay = np.asarray(assay, dtype=np.float64)
    return np.log2(assay + pseudo_count) if apply_log else assay      # prpsForCategoricalUV.R:87-98


def prps_categorical(assay, bio, uv, uv_variable, min_sample_for_prps=3, apply_log=True, pseudo_count=1.0):
    bio = np.array([str(x).replace("_", "-") for x in bio])           # :82-85
    uv = np.array([str(x).replace("_", "-") for x in uv])
    expre = _expre(assay, apply_log, pseudo_count)
    rows, cols = np.unique(bio), np.unique(uv)                        # table(): sorted levels (:100-103)
    tab = np.zeros((rows.size, cols.size), dtype=int)
    for b, u in zip(bio, uv):
        tab[np.searchsorted(rows, b), np.searchsorted(cols, u)] += 1
    bio_dist = (tab >= min_sample_for_prps).sum(axis=1)               # :104
    sel = bio_dist > 1                                                # :105
    if sel.sum() == 0:
        raise ValueError("There are not enough samples to create PRPS across the batches.")
    mats, names = [], []
    for y in np.nonzero(sel)[0]:                                      # :123-139
        for z in cols[tab[y] >= min_sample_for_prps]:
            index_sample = (bio == rows[y]) & (uv == z)
            mats.append(expre[:, index_sample].mean(axis=1))          # rowMeans
            names.append("%s||%s" % (rows[y], uv_variable))
    return np.stack(mats, axis=1), names


def prps_continuous(assay, uv_values, bio, batch, min_sample_for_prps=3, min_sample_per_batch=10, apply_log=True,
                    pseudo_count=1.0):
    expre = _expre(assay, apply_log, pseudo_count)
    uv_values = np.asarray(uv_values, dtype=np.float64)
    s_order = np.arange(len(bio))                                     # :166
    bio_batch = np.array(["%s_%s" % (bt, b) for bt, b in zip(batch, bio)])   # :171
    keys, n = np.unique(bio_batch, return_counts=True)
    keep = np.isin(bio_batch, keys[n >= min_sample_per_batch])        # add_count + filter (:172-173)
    s_order, uvk, bbk = s_order[keep], uv_values[keep], bio_batch[keep]

    def sliced(order):                                                # arrange(...) %>% group_by(bio.batch) %>% slice(1:min)
        so, bb = s_order[order], bbk[order]
        out_s, out_b = [], []
        for key in np.unique(bb):
            idx = np.nonzero(bb == key)[0][:min_sample_for_prps]
            out_s += list(so[idx])
            out_b += [key] * len(idx)
        return np.array(out_s), np.array(out_b)

    top_s, top_b = sliced(np.argsort(-uvk, kind="stable"))            # :174-177
    bot_s, _ = sliced(np.argsort(uvk, kind="stable"))                 # :178-181
    mats, names = [], []
    for y in range(0, len(bot_s), min_sample_for_prps):               # :190-201
        index = slice(y, y + min_sample_for_prps)
        mats.append(expre[:, bot_s[index]].mean(axis=1))
        mats.append(expre[:, top_s[index]].mean(axis=1))
        names += ["%s-LS" % top_b[y]] * 2
    return np.stack(mats, axis=1), names
