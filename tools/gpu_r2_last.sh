#!/bin/bash
# round 2, last call: GPU suite + default bench line of the final tree
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short > $O/r2l_test.log 2>&1
echo "exit $?" >> $O/r2l_test.log
grep -h "passed\|failed\|^exit" $O/r2l_test.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2l_test.log | cut -c1-300 | head -20
timeout -s USR1 -k 15 200 python bench.py --no-cpu-baseline > $O/r2l_bench.json 2> $O/r2l_bench.err
python - <<'PY'
import json
d = json.loads([x for x in open("gpurun_out/r2l_bench.json") if x.startswith("{")][-1])
print(round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["gpu_launches"], d["clocks"])
PY
