#!/usr/bin/env python
"""Timings of the steps either side of the path at the C3 size: K0 PRPS construction and the computePCA
fast path on the resident normalised assay.  One JSON line; device times from the library's CUDA events.

    python scripts/bench_aux.py [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--nb-pcs", type=int, default=10)
    args = ap.parse_args()
    from ruviii_b200 import _cabi, api
    from ruviii_b200.synth import make_config
    p = make_config("C3")
    m1, n, n_ps = p.meta["m1"], p.meta["n"], p.meta["n_ps"]
    eng = _cabi.Engine(devices=[0])
    groups = api.group_codes(list(p.sample_names) + list(p.replicate_names))
    prm = _cabi.make_params()
    Y = np.ascontiguousarray(p.Y)
    eng.plan(prm, m1, n, n_ps, groups, p.ctl, [p.k])
    eng.upload(Y, p.replicate_data)
    eng.run()
    out = {"workload": "C3: %d x %d, k=%d" % (m1, n, p.k)}
    # computePCA fast path on the resident newY (slot 0): no PCIe traffic except u, d, v coming back
    stages, wall = [], []
    for i in range(args.steps + 1):
        t0 = time.perf_counter()
        res = eng.pca(None, k=args.nb_pcs, slot=0)
        wall.append((time.perf_counter() - t0) * 1e3)
        stages.append(res["stage_ms"])
    st = {k: float(np.median([s[k] for s in stages[1:]])) for k in stages[0]}
    gram_flops = float(m1) * (m1 + 1) * n
    out["pca_resident"] = {"nb_pcs": args.nb_pcs, "stage_ms": st, "wall_ms": float(np.median(wall[1:])),
                           "gram_tflops": gram_flops / (st["gram"] * 1e-3) / 1e12,
                           "center_gbs": 8.0 * m1 * n * 3 / (st["center"] * 1e-3) / 1e9,
                           "d": [float(x) for x in res["d"][:4]]}
    # K11 gene statistics on the resident newY: ANOVA F over 73 batches and over 5 biologies, Pearson r with a covariate
    batch = ((np.arange(m1) // 5) % 73).astype(np.int32)
    bio = (np.arange(m1) % 5).astype(np.int32)
    cov = np.random.default_rng(2).gamma(5.0, 1.0, m1)
    gs = {}
    for name, kw in (("anova_73_batches", dict(groups=batch)), ("anova_5_biologies", dict(groups=bio)), ("pearson", dict(covariate=cov)),
                     ("anova+pearson", dict(groups=batch, covariate=cov))):
        ms = [eng.gene_stats(None, slot=0, **kw)["ms"] for _ in range(args.steps + 1)][1:]
        gs[name] = {"ms": float(np.median(ms)), "gbs": 8.0 * m1 * n / (float(np.median(ms)) * 1e-3) / 1e9}
    ms = [eng.gene_spearman(None, covariate=cov, slot=0)["ms"] for _ in range(3)][1:]
    gs["spearman"] = {"ms": float(np.median(ms)), "note": "K11b: one CTA per gene, shared-memory bitonic sort of 11,000 samples"}
    out["gene_stats_resident"] = gs
    # K0 PRPS: 1,200 pseudo-samples of 3 samples each, host matrix in, host PS out (only member rows are uploaded)
    rng = np.random.default_rng(1)
    sets = [np.sort(rng.choice(m1, 3, replace=False)).astype(np.int32) for _ in range(n_ps)]
    t = []
    for i in range(args.steps + 1):
        t0 = time.perf_counter()
        ps = eng.prps(Y, sets)
        t.append((time.perf_counter() - t0) * 1e3)
    out["prps_host_to_host"] = {"n_ps": n_ps, "wall_ms": float(np.median(t[1:])), "distinct_rows": int(len(set(np.concatenate(sets))))}
    ref = Y[sets[5]].mean(axis=0)
    out["prps_check_rel"] = float(np.linalg.norm(ps[5] - ref) / np.linalg.norm(ref))
    # fused K0: PS rows built on the device inside the resident pass
    eng.set_prps_sets(sets)
    eng.plan(prm, m1, n, n_ps, groups, p.ctl, [p.k])
    eng.upload(Y, None)
    info = eng.run()
    out["fused_prps_first_pass_ms"] = {"ms_log(K0 + log)": info.ms_log, "ms_total": info.ms_total}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
