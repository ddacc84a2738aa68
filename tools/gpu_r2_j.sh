#!/bin/bash
# round 2, GPU call J: same-box A/B of the pool-forward / stem-tail variants; 8-channel direct stem (teacher) tests + benches
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_kernels.py tests/test_net_parity.py -q -m gpu --tb=short -x -k "stem or birdview or pool or bn_forward or taps or full_size" > $O/r2j_test.log 2>&1
echo "exit $?" >> $O/r2j_test.log
for v in "0 0" "1 0" "0 1" "1 1" "0 0"; do
  set -- $v
  LBC_POOL_FWD=$1 LBC_STEM_TAIL=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2j_ab_$1$2.json 2> $O/r2j_ab_$1$2.err
  python - <<PY
import json
try:
    d = json.load(open("$O/r2j_ab_$1$2.json")); print("pool_fwd=$1 stem_tail=$2", round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()})
except Exception as ex:
    print("failed", ex); print(open("$O/r2j_ab_$1$2.err").read()[-600:])
PY
done
for w in config3 config5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r2j_bench_$w.json 2> $O/r2j_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("$O/r2j_bench_$w.json")); print("$w", round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, d["last_loss"])
except Exception as ex:
    print("failed", ex); print(open("$O/r2j_bench_$w.err").read()[-800:])
PY
done
grep -h "passed\|failed" $O/r2j_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2j_test.log | cut -c1-300 | head -20
