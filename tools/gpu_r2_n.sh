#!/bin/bash
# round 2, GPU call N: dlogits staged with cp.async in the head kernels, two-row writer of the space-to-depth stem operand
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > $O/r2n_test.log 2>&1
echo "exit $?" >> $O/r2n_test.log
grep -h "passed\|failed" $O/r2n_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2n_test.log | cut -c1-300 | head -20
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
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2n_bench_$i.json 2> $O/r2n_bench_$i.err
  show $O/r2n_bench_$i.json
done
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2n_ncu.log 2>&1
python tools/ncu_table.py launches $O/r2n_launches.csv | head -45
