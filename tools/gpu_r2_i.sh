#!/bin/bash
# round 2, GPU call I: suite + benches + launch list after the stem-tail fusion / pool forward / finalize unroll
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > $O/r2i_test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/r2i_test_gpu.log
for w in config2 config3 config5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r2i_bench_$w.json 2> $O/r2i_bench_$w.err
done
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/r2i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2i_launches.log 2>&1
grep -h "passed\|failed" $O/r2i_test_gpu.log | tail -2
grep -h "^FAILED\|^ERROR" $O/r2i_test_gpu.log | cut -c1-200 | head -20
grep -h "^E  " $O/r2i_test_gpu.log | cut -c1-300 | head -10
python - <<PY
import json
for f in ("config2", "config3", "config5"):
    try:
        d = json.load(open("$O/r2i_bench_%s.json" % f)); print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, "frac", round(d["roofline"]["frac"], 3), "loss", d["last_loss"])
    except Exception as ex:
        print(f, "failed", ex); print(open("$O/r2i_bench_%s.err" % f).read()[-800:])
PY
python tools/ncu_table.py launches $O/r2i_launches.csv 2>/dev/null | head -32
