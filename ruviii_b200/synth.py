"""Deterministic synthetic TCGA-shaped inputs for the RUV-III hot path.

log2-space model  Y = 1 mu' + B beta + W* alpha* + eps  with batch-structured unwanted factors, PRPS rows built
the way ``prpsForCategoricalUV`` does (prpsForCategoricalUV.R:123-139: a pseudo-sample is the mean of the
samples of one biology inside one batch; all pseudo-samples of one biology form one replicate set) and
negative-control genes drawn from the genes without a biological effect.

Used by tests and bench.py on the GPU box, so it reads nothing outside the repo.  numpy only.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BASE_SEED = 20261019

# name -> (m1 samples, n genes, G PRPS groups, reps per group, control genes, k)  [BASELINE.json configs]
CONFIGS = {
    "C1": dict(m1=1200, n=20000, groups=60, rep=3, n_ctl=1000, k=5, seed=BASE_SEED + 1),
    "C2": dict(m1=1200, n=20000, groups=60, rep=3, n_ctl=1000, k=5, seed=BASE_SEED + 1),   # same bytes as C1
    "C3": dict(m1=11000, n=20000, groups=400, rep=3, n_ctl=1000, k=10, seed=BASE_SEED + 3),
    "C4": dict(m1=11000, n=20000, groups=400, rep=3, n_ctl=1000, k=10, seed=BASE_SEED + 3),  # k = 1..20 on C3 data
    "C5p": dict(m1=60000, n=20000, groups=2000, rep=3, n_ctl=1000, k=10, seed=BASE_SEED + 5),
}


@dataclass
class Problem:
    """One synthetic RUV-III problem in the reference's own vocabulary."""
    assay: np.ndarray            # n x m1 (genes x samples), log2 scale unless ``counts`` was requested
    sample_names: list
    replicate_data: np.ndarray   # n_ps x n  (normalise.R:143)
    replicate_names: list        # one PRPS set id per pseudo-sample row
    ctl: np.ndarray              # bool[n]
    k: int
    meta: dict = field(default_factory=dict)

    @property
    def Y(self):
        """m1 x n row-major view of the assay (zero-copy when the assay is Fortran-ordered)."""
        return self.assay.T


def make_problem(m1, n, groups, rep=3, n_ctl=1000, k=5, seed=BASE_SEED, q=24, n_bio=5,
                 per_cell=3, counts=False, dtype=np.float64):
    """Generate one problem.  ``counts=True`` returns round(2^Y - 1) so that ``apply.log = TRUE`` can be tested."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_batches = max(2, min(max(8, m1 // 150), m1 // (n_bio * per_cell)))   # every (bio, batch) cell >= per_cell
    bio = np.arange(m1) % n_bio
    batch = (np.arange(m1) // n_bio) % n_batches            # round-robin: every (bio, batch) cell is populated
    mu_g = np.clip(rng.normal(8.0, 2.0, n), 0.0, 18.0)
    has_bio = rng.random(n) < 0.20
    beta = np.where(has_bio[None, :], rng.normal(0.0, 1.0, (n_bio, n)), 0.0)
    s = 1.5 * 0.85 ** np.arange(q)
    alpha_star = rng.normal(0.0, 1.0, (q, n)) * s[:, None]
    w_batch = rng.normal(0.0, 1.0, (n_batches, q))
    W_star = w_batch[batch] + rng.normal(0.0, 0.3, (m1, q))
    # samples x genes, C-ordered  ==  genes x samples Fortran-ordered (the R assay layout)
    Y = rng.normal(0.0, 0.4, (m1, n))
    Y += mu_g[None, :]
    Y += beta[bio]
    Y += W_star @ alpha_star
    null_genes = np.nonzero(~has_bio)[0]
    ctl = np.zeros(n, dtype=bool)
    ctl[rng.choice(null_genes, size=min(n_ctl, null_genes.size), replace=False)] = True

    # PRPS: group g belongs to biology g % n_bio; its `rep` pseudo-samples come from distinct batches
    cell_members = {}
    for i in range(m1):
        cell_members.setdefault((int(bio[i]), int(batch[i])), []).append(i)
    PS = np.empty((groups * rep, n), dtype=np.float64)
    rep_names = []
    for g in range(groups):
        b = g % n_bio
        cand = [bt for bt in range(n_batches) if len(cell_members.get((b, bt), ())) >= per_cell]
        if len(cand) < rep:
            raise ValueError("not enough populated batches for the requested PRPS design")
        chosen = rng.choice(cand, size=rep, replace=False)
        for j, bt in enumerate(chosen):
            members = rng.choice(cell_members[(b, int(bt))], size=per_cell, replace=False)
            PS[g * rep + j] = Y[members].mean(axis=0)
            rep_names.append("bio%d||set%04d" % (b, g))
    if counts:
        Y = np.round(np.exp2(Y) - 1.0).clip(min=0.0)
    assay = Y.T                                              # n x m1 view, Fortran-contiguous
    if dtype != np.float64:
        raise ValueError("the path is FP64 end to end")
    names = ["S%06d" % i for i in range(m1)]
    return Problem(assay=assay, sample_names=names, replicate_data=PS, replicate_names=rep_names,
                   ctl=ctl, k=k,
                   meta=dict(m1=m1, n=n, groups=groups, rep=rep, n_ps=groups * rep, n_ctl=int(ctl.sum()),
                             seed=seed, q=q, n_batches=n_batches))


def make_config(name, scale_rows=1.0):
    """One of BASELINE.json's configurations; ``scale_rows`` < 1 shrinks samples and groups for bounded CPU samples."""
    c = dict(CONFIGS[name])
    if scale_rows != 1.0:
        c["m1"] = max(50, int(round(c["m1"] * scale_rows)))
        c["groups"] = max(4, int(round(c["groups"] * scale_rows)))
    return make_problem(**c)


def make_dense_design(m, n, group_size=3, n_ctl=1000, k=10, seed=BASE_SEED + 30, q=24, cols=None, block=1024):
    """Dense replicate design: every row belongs to a replicate group of ``group_size``
    technical replicates (last group takes the remainder, minimum 2), no extra PRPS rows.  Returns
    (Y m x n row-major, group_ids int32[m], ctl bool[n]).

    Row factors come from ``seed``; gene columns are generated in blocks of ``block`` genes, block b from
    ``seed + 1 + b``, so ``cols = (g0, g1)`` returns exactly the columns [g0, g1) of the full matrix without
    generating the rest (one rank of a gene-sharded bench holds only its slab); ctl is restricted likewise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_groups = max(1, m // group_size)
    gid = np.minimum(np.arange(m) // group_size, n_groups - 1).astype(np.int32)
    s = 1.5 * 0.85 ** np.arange(q)
    group_factors = rng.normal(0.0, 1.0, (n_groups, 6))[gid]          # per-group biology
    W_star = rng.normal(0.0, 1.0, (m, q))                             # unwanted variation differs between replicates
    ctl = np.zeros(n, dtype=bool)
    ctl[rng.choice(n, size=min(n_ctl, n), replace=False)] = True
    g0, g1 = (0, n) if cols is None else cols
    Y = np.empty((m, g1 - g0), dtype=np.float64)
    for b in range(g0 // block, (g1 + block - 1) // block):
        lo, hi = b * block, min((b + 1) * block, n)
        rb = np.random.Generator(np.random.PCG64(seed + 1 + b))
        w = hi - lo
        mu_g = np.clip(rb.normal(8.0, 2.0, w), 0.0, 18.0)
        alpha_star = rb.normal(0.0, 1.0, (q, w)) * s[:, None]
        loadings = rb.normal(0.0, 0.5, (6, w))
        blk = rb.normal(0.0, 0.4, (m, w))
        blk += mu_g[None, :]
        blk += group_factors @ loadings
        blk += W_star @ alpha_star
        a, z = max(lo, g0), min(hi, g1)
        Y[:, a - g0:z - g0] = blk[:, a - lo:z - lo]
    return Y, gid, ctl[g0:g1]
