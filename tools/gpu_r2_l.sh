#!/bin/bash
# round 2, GPU call L: space-to-depth stem (A/B on one box: LBC_PAIR=7 is the 7-row layout), mma.sync head kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_kernels.py tests/test_ops.py tests/test_net_parity.py -q -m gpu --tb=short -x \
    -k "stem or head or taps or golden or full_size or birdview or B32 or phase" > $O/r2l_test.log 2>&1
echo "exit $?" >> $O/r2l_test.log
grep -h "passed\|failed" $O/r2l_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2l_test.log | cut -c1-300 | head -20
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fp32frames", round(d["e2e"]["fp32_frames"]["value"]),
          "frac", round(d["roofline"]["frac"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
for v in "7 1" "15 1" "15 0" "7 1" "15 1"; do
  set -- $v
  LBC_PAIR=$1 LBC_STEM_S2D_CFG=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2l_ab_$1_$2.json 2> $O/r2l_ab_$1_$2.err
  show $O/r2l_ab_$1_$2.json
done
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/r2l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2l_ncu.log 2>&1
grep -c . $O/r2l_launches.csv
