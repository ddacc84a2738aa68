#!/bin/bash
# round 2, GPU call Y: nine-tap weight gradient of the 64 -> 64 convolutions (wgrad9_c64_kernel): per-op parity, whole net, A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_ops.py tests/test_kernels.py -q -m gpu --tb=short -k "wgrad9 or copy_channels" > $O/r2y_test_ops.log 2>&1
echo "exit $?" >> $O/r2y_test_ops.log
grep -h "passed\|failed\|^exit" $O/r2y_test_ops.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2y_test_ops.log | cut -c1-300 | head -30
LBC_PAIR=127 timeout 400 python -m pytest tests/test_net_parity.py tests/test_training_entry.py -q -m gpu --tb=short -k "bf16 or full_size or schedule" > $O/r2y_test_net.log 2>&1
echo "exit $?" >> $O/r2y_test_net.log
grep -h "passed\|failed\|^exit" $O/r2y_test_net.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2y_test_net.log | cut -c1-300 | head -20
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
for v in 63 127 63 127; do
  LBC_PAIR=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2y_ab_pair${v}_$RANDOM.json 2> $O/r2y_ab.err
  show $(ls -t $O/r2y_ab_pair${v}_*.json | head -1)
done
