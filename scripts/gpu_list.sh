#!/bin/bash
# launch list only: gpurun --timeout 600 -- 'bash scripts/gpu_list.sh <tag>'
TAG=${1:-r2list}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra"
timeout -s KILL 300 $CMD > $OUT/plain.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "ncu list exit $?"
python - "$OUT/launches.csv" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict(); cnt = collections.Counter()
for r in rows[hi + 2:]:
    if len(r) <= vi: continue
    nm = r[ki][:70]; v = float(r[vi].replace(",", "")) / 1e3
    agg[nm] = agg.get(nm, 0.0) + v; cnt[nm] += 1
for nm, v in sorted(agg.items(), key=lambda kv: -kv[1])[:16]:
    print("%9.1f us x%d  %7.1f  %s" % (v, cnt[nm], v / cnt[nm], nm))
PY
