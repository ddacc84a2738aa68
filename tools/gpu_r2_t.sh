#!/bin/bash
# round 2, GPU call T: timing experiments (results of the SKIP runs are garbage by construction): what the small finalize
# kernels cost in the pipeline, what the side-stream weight gradients still cost, 4-stage pair weight gradient
mkdir -p gpurun_out
O=gpurun_out
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-800:])
PY
}
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2t_$name.json 2> $O/r2t_$name.err
  show $O/r2t_$name.json
}
run base A=1
run skipfin LBC_EXPERIMENT_SKIP_FINALIZE=1
run skipwgrad LBC_EXPERIMENT_SKIP_WGRAD=1
run w3s4 LBC_W3_STAGES=4
run ovl2 LBC_WGRAD_OVERLAP=2
run base2 A=1
