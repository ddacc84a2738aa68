#!/bin/bash
# round 2, GPU call G: suite + bench after the BN finalize fusion, per-CTA statistics of the N=64 kernels, fast head_dlogits
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > $O/r2g_test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/r2g_test_gpu.log
for w in config2 config3 config5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r2g_bench_$w.json 2> $O/r2g_bench_$w.err
done
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2g_launches.log 2>&1
grep -h "passed\|failed" $O/r2g_test_gpu.log | tail -2
grep -h "^FAILED\|^ERROR" $O/r2g_test_gpu.log | cut -c1-200 | head -20
grep -h "^E  " $O/r2g_test_gpu.log | cut -c1-300 | head -10
python - <<PY
import json
for f in ("config2", "config3", "config5"):
    try:
        d = json.load(open("$O/r2g_bench_%s.json" % f)); print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, "frac", round(d["roofline"]["frac"], 3), "loss", d["last_loss"])
    except Exception as ex:
        print(f, "failed", ex); print(open("$O/r2g_bench_%s.err" % f).read()[-800:])
PY
python tools/ncu_table.py launches $O/r2g_launches.csv 2>/dev/null | head -34
