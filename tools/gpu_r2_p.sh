#!/bin/bash
# round 2, GPU call P: residual add fused with the next BatchNorm's reduce pass (A/B on one box), inference-graph test
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > $O/r2p_test.log 2>&1
echo "exit $?" >> $O/r2p_test.log
grep -h "passed\|failed" $O/r2p_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2p_test.log | cut -c1-300 | head -20
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "frac", round(d["roofline"]["frac"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
for v in 0 1 0 1; do
  LBC_RESID_FUSE=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2p_ab_$v.json 2> $O/r2p_ab_$v.err
  show $O/r2p_ab_$v.json
done
