#!/bin/bash
# round 2, GPU call U: timing experiments (SKIP runs produce garbage by construction): forward / backward finalize kernels
# separately, and the bn1 reduce pass a data-gradient epilogue fusion would remove
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
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2u_$name.json 2> $O/r2u_$name.err
  show $O/r2u_$name.json
}
run base A=1
run skipfin_fwd LBC_EXPERIMENT_SKIP_FINALIZE=1
run skipfin_bwd LBC_EXPERIMENT_SKIP_FINALIZE=2
run skip_ownreduce LBC_EXPERIMENT_SKIP_OWNREDUCE=1
run serial_base LBC_WGRAD_OVERLAP=0
run serial_skip_ownreduce LBC_WGRAD_OVERLAP=0 LBC_EXPERIMENT_SKIP_OWNREDUCE=1
