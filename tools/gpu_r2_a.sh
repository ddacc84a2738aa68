#!/bin/bash
# round 2, GPU call A: full GPU suite with the new per-kernel parity tests (no -x: collect every failure), the
# EXPERIMENTAL candidates one by one (bench A/B on this box), the tensor-core accumulation probe.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -s > $O/r2a_test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/r2a_test_gpu.log
LBC_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_net_parity.py -q -m gpu -k experimental --tb=short > $O/r2a_test_exp.log 2>&1
echo "experimental exit $?" >> $O/r2a_test_exp.log
for e in 0 1 2 4 8 15; do
  LBC_EXPERIMENTAL=$e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2a_bench_exp$e.json 2> $O/r2a_bench_exp$e.err
done
grep -h "passed\|failed\|error" $O/r2a_test_gpu.log | tail -3
grep -h "FAILED\|ERROR" $O/r2a_test_gpu.log | head -40
tail -2 $O/r2a_test_exp.log
for e in 0 1 2 4 8 15; do python - <<PY
import json
try:
    d=json.load(open('$O/r2a_bench_exp$e.json')); print('exp$e', round(d['ms_per_step'],3), 'ms', round(d['value']), 'img/s', {k: round(v['ms_per_step'],3) for k,v in d['roofline']['per_category'].items()})
except Exception as ex: print('exp$e failed', ex)
PY
done
