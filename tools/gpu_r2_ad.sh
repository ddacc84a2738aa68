#!/bin/bash
# round 2, GPU call AD: layer 1 (64 -> 64 3x3/s1) on the shared-row CTA-pair kernel with 64-wide N tiles (LBC_PAIR bit 7)
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_ops.py tests/test_kernels.py -q -m gpu --tb=short -k "row64 or (conv_epilogue_statistics)" > $O/r2ad_test_ops.log 2>&1
echo "exit $?" >> $O/r2ad_test_ops.log
grep -h "passed\|failed\|^exit" $O/r2ad_test_ops.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2ad_test_ops.log | cut -c1-300 | head -20
LBC_PAIR=255 timeout 400 python -m pytest tests/test_net_parity.py -q -m gpu --tb=short -k "bf16 or schedule or full_size" > $O/r2ad_test_net.log 2>&1
echo "exit $?" >> $O/r2ad_test_net.log
grep -h "passed\|failed\|^exit" $O/r2ad_test_net.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2ad_test_net.log | cut -c1-300 | head -10
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3),
          {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_category"].items()}, d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
for v in 127 255 127 255; do
  LBC_PAIR=$v timeout -s USR1 -k 15 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2ad_pair${v}_$RANDOM.json 2> $O/r2ad.err
  show $(ls -t $O/r2ad_pair${v}_*.json | head -1)
done
