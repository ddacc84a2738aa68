#!/bin/bash
# round 2, GPU call AC: split reduction dealt to four thread groups, lean MMA issue loop in the nine-tap weight gradient
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_ops.py -q -m gpu --tb=short -k "tcgen05_conv" > $O/r2ac_test_ops.log 2>&1
echo "exit $?" >> $O/r2ac_test_ops.log
grep -h "passed\|failed\|^exit" $O/r2ac_test_ops.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2ac_test_ops.log | cut -c1-300 | head -20
timeout 400 python -m pytest tests/test_net_parity.py -q -m gpu --tb=short -k "bf16 or schedule or full_size" > $O/r2ac_test_net.log 2>&1
echo "exit $?" >> $O/r2ac_test_net.log
grep -h "passed\|failed\|^exit" $O/r2ac_test_net.log | tail -3
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3),
          {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
for i in 1 2; do
  timeout -s USR1 -k 15 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2ac_bench_$i.json 2> $O/r2ac_bench_$i.err; show $O/r2ac_bench_$i.json
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'wgrad' -c 120 --csv --log-file $O/r2ac_wgrad_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2ac_launches.log 2>&1
python tools/ncu_table.py launches $O/r2ac_wgrad_launches.csv | head -12
