#!/bin/bash
# round 2, final evidence call: full GPU suite, smoke, the three workloads, launch list, ncu --set full of the GEMM families
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short > $O/r2f_test.log 2>&1
echo "exit $?" >> $O/r2f_test.log
grep -h "passed\|failed\|^exit" $O/r2f_test.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2f_test.log | cut -c1-300 | head -20
timeout 300 python __graft_entry__.py smoke > $O/r2f_smoke.log 2>&1; tail -5 $O/r2f_smoke.log
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3),
          {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, d["clocks"], d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
timeout 400 python bench.py > $O/r2f_bench_default.json 2> $O/r2f_bench_default.err; show $O/r2f_bench_default.json
for w in config2 config3 config5; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $w > $O/r2f_bench_$w.json 2> $O/r2f_bench_$w.err
  show $O/r2f_bench_$w.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/r2f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2f_launches.log 2>&1
python tools/ncu_table.py launches $O/r2f_launches.csv | head -45
timeout 400 ncu --set full --clock-control none \
    -k regex:'conv_gemm_kernel|conv_row_kernel|conv3x3_c64|wgrad|bn_apply_kernel|bn_bwd|stem_conv|stem_wgrad' \
    --launch-skip 200 --launch-count 60 -o $O/r2f_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r2f_full.log 2>&1
ncu -i $O/r2f_full.ncu-rep --page raw --csv > $O/r2f_full.csv 2>/dev/null
rm -f $O/r2f_full.ncu-rep
python tools/ncu_table.py full $O/r2f_full.csv | cut -c1-220 | head -70
