"""CPU ORACLE (test infrastructure, never shipped, never imported by ruviii_b200/) for computePCA's fast path
(R/computePCA.R:120-141): ``runSVD(t(assay), k, BSPARAM = bsparam(), center, scale)`` with the default ExactParam is
base ``svd`` of the centred / scaled matrix truncated to k (BiocSingular documentation; the package is not in the
container, so this is PARITY UNPINNED like the rest).  LAPACK dgesdd through numpy, as R uses."""
import numpy as np


def fast_pca(assay, k, apply_log=True, pseudo_count=1.0, center=True, scale=False):
    x = np.asarray(assay, dtype=np.float64)
    if apply_log:
        x = np.log2(x + pseudo_count)                            # computePCA.R:124-127
    x = x.T                                                      # transpose(temp.data): samples x genes
    if center:
        x = x - x.mean(axis=0, keepdims=True)
    if scale:                                                    # base::scale: sd with m - 1 (root mean square if not centred)
        x = x / np.sqrt((x ** 2).sum(axis=0, keepdims=True) / (x.shape[0] - 1))
    u, d, vt = np.linalg.svd(x, full_matrices=False)
    pct = np.round(d[:k] ** 2 / np.sum(d[:k] ** 2) * 100.0, 1)   # :137-141 (relative to the k computed values)
    return dict(u=u[:, :k], d=d[:k], v=vt[:k].T, variation=pct)
