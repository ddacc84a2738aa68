#!/bin/bash
# round 2, GPU call D: full suite (fp32tc reordered terms, config-1 plumbing), the three bench workloads, the reference arm
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -s > $O/r2d_test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/r2d_test_gpu.log
for w in config2 config3 config5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r2d_bench_$w.json 2> $O/r2d_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/r2d_bench_ref.json 2> $O/r2d_bench_ref.err
grep -h "passed\|failed" $O/r2d_test_gpu.log | tail -2
grep -h "^FAILED\|^ERROR" $O/r2d_test_gpu.log | cut -c1-200 | head -40
grep -h "^fp32tc\|config1 B=8\|taps\[fp32tc" $O/r2d_test_gpu.log | cut -c1-700 | head -40
python - <<PY
import json
for f in ("config2", "config3", "config5"):
    try:
        d = json.load(open("$O/r2d_bench_%s.json" % f)); print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "fp32frames", round(d["e2e"]["fp32_frames"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, "frac", round(d["roofline"]["frac"], 3), "loss", d["last_loss"])
    except Exception as ex:
        print(f, "failed", ex); print(open("$O/r2d_bench_%s.err" % f).read()[-800:])
try:
    d = json.load(open("$O/r2d_bench_ref.json")); print("ref", d["value"], d["config"], d["cpu_baseline"])
except Exception as ex:
    print("ref failed", ex); print(open("$O/r2d_bench_ref.err").read()[-800:])
PY
