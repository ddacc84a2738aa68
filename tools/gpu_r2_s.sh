#!/bin/bash
# round 2, GPU call S: new defaults (shared-row kernel on layers 2-3, weight gradients on the side stream, PDL): full GPU
# suite, bench lines of the three workloads, launch list, ncu --set full with source of the layer-1 / shared-row kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short -x > $O/r2s_test.log 2>&1
echo "exit $?" >> $O/r2s_test.log
grep -h "passed\|failed\|^exit" $O/r2s_test.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2s_test.log | cut -c1-300 | head -20
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
for w in config2 config3 config5; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $w > $O/r2s_bench_$w.json 2> $O/r2s_bench_$w.err
  show $O/r2s_bench_$w.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/r2s_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2s_launches.log 2>&1
python tools/ncu_table.py launches $O/r2s_launches.csv | head -50
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_c64|conv_row_kernel' \
    --launch-skip 3 --launch-count 5 -o $O/r2s_src -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r2s_src.log 2>&1
ls -la $O/r2s_src.ncu-rep
timeout 400 ncu --set full --clock-control none \
    -k regex:'conv_gemm_kernel|conv_row_kernel|conv3x3_c64|wgrad|bn_apply_kernel|bn_bwd|bn_finalize|ew_kernel' \
    --launch-skip 60 --launch-count 70 -o $O/r2s_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r2s_full.log 2>&1
ncu -i $O/r2s_full.ncu-rep --page raw --csv > $O/r2s_full.csv 2>/dev/null
ls -la $O/r2s_full.ncu-rep
rm -f $O/r2s_full.ncu-rep
python tools/ncu_table.py full $O/r2s_full.csv | cut -c1-220 | head -80
