"""Generates the golden fixtures in this directory from the CPU oracle (oracle/ruviii_oracle.py, literal flavour).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference ships no golden vectors and R is not installed in the build container, so
these files pin the *restatement*, not an R session.  Every matrix is stored exactly as R would hold it --
raw little-endian float64, column-major -- next to a JSON sidecar with dim/dimnames (for the Python tests) and a
base-R-readable sidecar (<case>.manifest.tsv, <case>.sample_names.txt, <case>.replicate_names.txt,
<case>.ctl_index_1based.txt), so that ONE command in an R session that has the reference installed settles the pin:

    Rscript tests/golden/replay.R            # replays every fixture through the reference's own ruvIII / fastruvIII /
                                             # ruvIIIMultipleK and exits non-zero above 1e-9 (rel. Frobenius)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ncg_oracle as NO            # noqa: E402
from oracle import pca_oracle as PCAO          # noqa: E402
from oracle import ruviii_oracle as O          # noqa: E402
from ruviii_b200.synth import make_problem     # noqa: E402


def put(case, name, a):
    a = np.asarray(a, dtype="<f8")
    a.T.ravel().tofile(os.path.join(HERE, "%s.%s.bin" % (case, name)))     # column-major bytes
    return list(a.shape)


def tiny():
    # hand-checkable: 4 samples, 2 PRPS sets x 2 pseudo-samples, 6 genes, 3 control genes, k = 1
    assay = np.array([[5.0, 6.0, 7.0, 8.0],
                      [4.0, 4.5, 6.0, 6.5],
                      [9.0, 8.0, 7.0, 6.0],
                      [3.0, 3.5, 3.0, 3.5],
                      [7.0, 7.0, 8.0, 8.0],
                      [2.0, 4.0, 6.0, 8.5]])
    PS = np.array([[5.5, 4.2, 8.5, 3.2, 7.0, 3.0],
                   [7.5, 6.3, 6.5, 3.3, 8.1, 7.2],
                   [5.0, 4.0, 9.1, 3.0, 7.2, 2.1],
                   [8.1, 6.4, 6.1, 3.6, 7.9, 8.4]])
    return dict(assay=assay, sample_names=["s1", "s2", "s3", "s4"], replicate_data=PS,
                replicate_names=["A", "A", "B", "B"], ctl=np.array([True, False, True, False, True, False]), k=1)


def mid():
    p = make_problem(m1=60, n=300, groups=7, rep=3, n_ctl=40, k=4, seed=20261019 + 100)
    return dict(assay=np.array(p.assay), sample_names=p.sample_names, replicate_data=p.replicate_data,
                replicate_names=p.replicate_names, ctl=p.ctl, k=4)


def main():
    for case, c in (("tiny", tiny()), ("mid", mid())):
        meta = dict(case=case, k=int(c["k"]), ks_multi=[1, 2, c["k"]] if case == "mid" else [1],
                    sample_names=list(c["sample_names"]), replicate_names=list(c["replicate_names"]),
                    ctl_index_0based=[int(i) for i in np.nonzero(c["ctl"])[0]], apply_log=False,
                    layout="raw little-endian float64, column-major (R)", dims={})
        meta["dims"]["assay"] = put(case, "assay", c["assay"])
        meta["dims"]["replicate_data"] = put(case, "replicate_data", c["replicate_data"])
        L = O.ruvIII_literal(c["assay"], c["sample_names"], c["replicate_data"], c["replicate_names"], c["ctl"],
                             c["k"], apply_log=False)
        meta["dims"]["ruvIII.newY"] = put(case, "ruvIII.newY", L["newY"])
        meta["dims"]["ruvIII.fullalpha"] = put(case, "ruvIII.fullalpha", L["fullalpha"])
        meta["dims"]["ruvIII.W"] = put(case, "ruvIII.W", L["W"])
        F = O.fastruvIII_literal(c["assay"], c["replicate_data"], c["replicate_names"], c["ctl"], c["k"],
                                 apply_log=False)
        meta["dims"]["fastruvIII.newY"] = put(case, "fastruvIII.newY", F["newY"])
        meta["dims"]["fastruvIII.fullalpha"] = put(case, "fastruvIII.fullalpha", F["fullalpha"])
        meta["dims"]["fastruvIII.W"] = put(case, "fastruvIII.W", F["W"])
        for k in meta["ks_multi"]:
            Lk = O.ruvIII_literal(c["assay"], c["sample_names"], c["replicate_data"], c["replicate_names"],
                                  c["ctl"], k, apply_log=False)
            meta["dims"]["ruvIIIMultipleK.k%d.newY" % k] = put(case, "ruvIIIMultipleK.k%d.newY" % k, Lk["newY"])
        # apply.log = TRUE on counts (ruvIII.R:88-89) for the mid case
        if case == "mid":
            counts = np.round(np.exp2(c["assay"]) - 1.0).clip(min=0.0)
            meta["dims"]["counts"] = put(case, "counts", counts)
            Lc = O.ruvIII_literal(counts, c["sample_names"], c["replicate_data"], c["replicate_names"], c["ctl"],
                                  c["k"], apply_log=True, pseudo_count=1.0)
            meta["dims"]["ruvIII.log.newY"] = put(case, "ruvIII.log.newY", Lc["newY"])
        # eta != NULL (ruv::RUV1 pre-step, ruvIII.R:116): a gene-wise covariate, intercept added by design.matrix
        if case == "mid":
            n = c["assay"].shape[0]
            eta = np.cos(np.arange(n) * 0.37)[None, :] + 0.1 * np.arange(n)[None, :] / n
            meta["dims"]["eta"] = put(case, "eta", eta.T)              # n x 1, the orientation R users pass
            Le = O.ruvIII_literal(c["assay"], c["sample_names"], c["replicate_data"], c["replicate_names"], c["ctl"],
                                  c["k"], apply_log=False, eta=eta, include_intercept=True)
            meta["dims"]["ruvIII.eta.newY"] = put(case, "ruvIII.eta.newY", Le["newY"])
        # the steps either side of the path whose reference semantics are BASE R: lm residuals, one-way ANOVA F,
        # Pearson / Spearman correlation, svd of the centred matrix -- replay.R checks these without any add-on package
        if case == "mid":
            m1 = c["assay"].shape[1]
            batch = ["b%d" % ((i * 7 + i // 5) % 5) for i in range(m1)]
            cov = np.sin(np.arange(m1) * 0.61) + 0.02 * np.arange(m1)
            cov[::9] = cov[0]                                           # ties for Spearman
            meta["side_batch"] = batch
            meta["dims"]["side.covariate"] = put(case, "side.covariate", cov[:, None])
            cd = {"batch": batch, "cov": list(cov)}
            meta["dims"]["side.lm_residuals"] = put(case, "side.lm_residuals", NO.lm_residuals(c["assay"], cd, ["batch", "cov"]))
            meta["dims"]["side.anovaF"] = put(case, "side.anovaF", NO.row_oneway_equalvar_statistic(c["assay"], np.asarray(batch))[:, None])
            meta["dims"]["side.pearson"] = put(case, "side.pearson", NO.correls_pearson(cov, c["assay"].T)[:, None])
            meta["dims"]["side.spearman"] = put(case, "side.spearman", NO.correls_spearman(cov, c["assay"].T)[:, None])
            sv = PCAO.fast_pca(c["assay"], min(c["assay"].shape), apply_log=False)
            meta["dims"]["side.svd_d"] = put(case, "side.svd_d", sv["d"][:, None])
            meta["dims"]["side.svd_rank10"] = put(case, "side.svd_rank10", (sv["u"][:, :10] * sv["d"][:10]) @ sv["v"][:, :10].T)
            with open(os.path.join(HERE, "%s.side_batch.txt" % case), "w") as f:
                f.write("\n".join(batch) + "\n")
        with open(os.path.join(HERE, "%s.json" % case), "w") as f:
            json.dump(meta, f, indent=1)
        # base-R-readable sidecars for replay.R
        with open(os.path.join(HERE, "%s.manifest.tsv" % case), "w") as f:
            f.write("name\tnrow\tncol\n")
            for name, dim in meta["dims"].items():
                f.write("%s\t%d\t%d\n" % (name, dim[0], dim[1]))
            f.write("k\t%d\t0\n" % meta["k"])
            for kv in meta["ks_multi"]:
                f.write("k_multi\t%d\t0\n" % kv)
        for key, vals in (("sample_names", meta["sample_names"]), ("replicate_names", meta["replicate_names"]),
                          ("ctl_index_1based", [i + 1 for i in meta["ctl_index_0based"]])):
            with open(os.path.join(HERE, "%s.%s.txt" % (case, key)), "w") as f:
                f.write("\n".join(str(v) for v in vals) + "\n")
        print(case, {k: v for k, v in meta["dims"].items()})


if __name__ == "__main__":
    main()
