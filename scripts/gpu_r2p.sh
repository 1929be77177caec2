#!/bin/bash
TAG=${1:-r2p}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests/test_pca.py tests/test_ncg.py tests/test_gpu_parity.py tests/test_prps.py -m gpu -q -x --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest.log
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - "$OUT/bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   ms/step %.3f  " % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
PY
