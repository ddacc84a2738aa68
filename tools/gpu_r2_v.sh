#!/bin/bash
# round 2, GPU call V: BatchNorm statistics finalised by the convolution's last CTA (lbc_bn_tail.h): parity, A/B, full suite
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_net_parity.py -q -m gpu --tb=short -x -k "statistics_tail or launch_schedule or taps_match" > $O/r2v_test_tail.log 2>&1
echo "exit $?" >> $O/r2v_test_tail.log
grep -h "passed\|failed\|^exit" $O/r2v_test_tail.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2v_test_tail.log | cut -c1-400 | head -20
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2v_$name.json 2> $O/r2v_$name.err
  show $O/r2v_$name.json
}
run tail0 LBC_BN_TAIL=0
run tail1 LBC_BN_TAIL=1
run tail0b LBC_BN_TAIL=0
run tail1b LBC_BN_TAIL=1
timeout 900 python -m pytest tests -q -m gpu --tb=short -x > $O/r2v_test_all.log 2>&1
echo "exit $?" >> $O/r2v_test_all.log
grep -h "passed\|failed\|^exit" $O/r2v_test_all.log | tail -3
grep -h "^FAILED\|^ERROR\|^E  " $O/r2v_test_all.log | cut -c1-400 | head -20
