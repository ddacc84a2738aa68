#!/bin/bash
# round 2, GPU call K: full GPU suite after removing the losing variants; per-CTA BN statistics rows in the GEMM epilogue;
# fixed-slot prefetcher (end-to-end stability: the bench is run three times); eval-forward latency probe
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > $O/r2k_test.log 2>&1
echo "exit $?" >> $O/r2k_test.log
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
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2k_bench_$i.json 2> $O/r2k_bench_$i.err
  show $O/r2k_bench_$i.json
done
for w in config3 config5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r2k_bench_$w.json 2> $O/r2k_bench_$w.err
  show $O/r2k_bench_$w.json
done
timeout 300 python tools/gpu_infer_probe.py > $O/r2k_infer.jsonl 2> $O/r2k_infer.err; cat $O/r2k_infer.jsonl; tail -3 $O/r2k_infer.err
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2k_ncu.log 2>&1
grep -c . $O/r2k_launches.csv
grep -h "passed\|failed" $O/r2k_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2k_test.log | cut -c1-300 | head -20
