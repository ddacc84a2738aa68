#!/bin/bash
# round 2, GPU call Q: launch schedules -- weight gradients on a side stream (LBC_WGRAD_OVERLAP) and programmatic dependent
# launch (LBC_PDL): result invariance test, then same-box A/B of bench.py
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_net_parity.py -q -m gpu --tb=short -x -k "launch_schedule" > $O/r2q_test_sched.log 2>&1
echo "exit $?" >> $O/r2q_test_sched.log
grep -h "passed\|failed\|^exit" $O/r2q_test_sched.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2q_test_sched.log | cut -c1-300 | head -20
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "frac", round(d["roofline"]["frac"], 3), d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
for v in "0 0" "2 0" "1 0" "0 1" "2 1"; do
  set -- $v
  LBC_WGRAD_OVERLAP=$1 LBC_PDL=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2q_ab_ovl$1_pdl$2.json 2> $O/r2q_ab_ovl$1_pdl$2.err
  show $O/r2q_ab_ovl$1_pdl$2.json
done
# the per-kernel and whole-net tests under the new schedule (every kernel launched with the PDL attribute; graph capture of PDL launches)
LBC_WGRAD_OVERLAP=2 LBC_PDL=1 timeout 1200 python -m pytest tests -q -m gpu --tb=short -x > $O/r2q_test_all_sched.log 2>&1
echo "exit $?" >> $O/r2q_test_all_sched.log
grep -h "passed\|failed\|^exit" $O/r2q_test_all_sched.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2q_test_all_sched.log | cut -c1-300 | head -20
