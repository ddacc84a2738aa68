#!/bin/bash
# round 2, last 2-GPU check of the final tree: one data-parallel bench run
mkdir -p gpurun_out
timeout -s USR1 -k 15 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l2_n2.json 2> gpurun_out/r2l2_n2.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open("gpurun_out/r2l2_n2.json") if x.startswith("{")][-1])
    print(d["n_gpus"], "gpus", round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["last_loss"])
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r2l2_n2.err").read()[-1500:])
PY
