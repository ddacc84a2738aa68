#!/bin/bash
# round 2, GPU call AA: ring size of the side-stream gradient buffers (same-box A/B), defaults = LBC_PAIR 127
mkdir -p gpurun_out
O=gpurun_out
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
  env "$@" timeout -s USR1 -k 15 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2aa_$name.json 2> $O/r2aa_$name.err
  show $O/r2aa_$name.json
}
run ring4 LBC_RING=4
run ring2 LBC_RING=2
run ring8 LBC_RING=8
run ring4b LBC_RING=4
run ring8b LBC_RING=8
run ovl2 LBC_WGRAD_OVERLAP=2
