#!/bin/bash
# round 2, GPU call C: full GPU suite after the BatchNorm-backward fix + fp32tc without format mixing; smoke; bench bf16 + fp32tc
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -s > $O/r2c_test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/r2c_test_gpu.log
timeout 600 python __graft_entry__.py smoke > $O/r2c_smoke.log 2>&1
echo "exit $?" >> $O/r2c_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2c_bench.json 2> $O/r2c_bench.err
timeout 400 python bench.py --precision fp32tc --batch 64 --steps 3 --warmup 1 --no-cpu-baseline > $O/r2c_bench_tc64.json 2> $O/r2c_bench_tc64.err
grep -h "passed\|failed" $O/r2c_test_gpu.log | tail -2
grep -h "^FAILED\|^ERROR" $O/r2c_test_gpu.log | cut -c1-200 | head -40
grep -h "^fp32tc\|^taps\|tcgen05 fp32 acc\|bf16 (fast\|median error" $O/r2c_test_gpu.log | cut -c1-900 | head -70
tail -6 $O/r2c_smoke.log | cut -c1-300
python - <<PY
import json
for f in ("r2c_bench", "r2c_bench_tc64"):
    try:
        d = json.load(open("$O/%s.json" % f)); print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, "frac", round(d["roofline"]["frac"], 3))
    except Exception as ex:
        print(f, "failed", ex); print(open("$O/%s.err" % f).read()[-600:])
PY
