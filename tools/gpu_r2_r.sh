#!/bin/bash
# round 2, GPU call R: shared-row CTA-pair kernel (conv_row_kernel<128>) -- per-op parity, statistics, whole net, A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_ops.py tests/test_kernels.py -q -m gpu --tb=short -k "rowk or row" > $O/r2r_test_ops.log 2>&1
echo "exit $?" >> $O/r2r_test_ops.log
grep -h "passed\|failed\|^exit" $O/r2r_test_ops.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2r_test_ops.log | cut -c1-300 | head -30
LBC_PAIR=63 timeout 400 python -m pytest tests/test_net_parity.py tests/test_training_entry.py -q -m gpu --tb=short -k "bf16 or full_size or schedule" > $O/r2r_test_net.log 2>&1
echo "exit $?" >> $O/r2r_test_net.log
grep -h "passed\|failed\|^exit" $O/r2r_test_net.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2r_test_net.log | cut -c1-300 | head -20
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
for v in "15 0 0" "31 0 0" "63 0 0" "15 1 1" "31 1 1" "63 1 1" "63 2 1"; do
  set -- $v
  LBC_PAIR=$1 LBC_WGRAD_OVERLAP=$2 LBC_PDL=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2r_ab_pair$1_ovl$2_pdl$3.json 2> $O/r2r_ab_pair$1_ovl$2_pdl$3.err
  show $O/r2r_ab_pair$1_ovl$2_pdl$3.json
done
