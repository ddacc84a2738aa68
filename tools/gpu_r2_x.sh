#!/bin/bash
# round 2, GPU call X: weight packs on the side stream during the stem, vectorised speed-fusion / slice kernels: suite + A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short -x > $O/r2x_test.log 2>&1
echo "exit $?" >> $O/r2x_test.log
grep -h "passed\|failed\|^exit" $O/r2x_test.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2x_test.log | cut -c1-400 | head -20
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2x_$name.json 2> $O/r2x_$name.err
  show $O/r2x_$name.json
}
run packside0 LBC_PACK_SIDE=0
run packside1 LBC_PACK_SIDE=1
run packside0b LBC_PACK_SIDE=0
run packside1b LBC_PACK_SIDE=1
