#!/bin/bash
# round 2, GPU call W (2 GPUs): data-parallel step under the new schedule (side-stream weight gradients, PDL, shared-row kernel):
# single all-reduce and bucketed+overlapped; rank-consistency of the weights after the run is checked by bench.py's loss line
mkdir -p gpurun_out
O=gpurun_out
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    l = [x for x in open(f) if x.startswith("{")][-1]
    d = json.loads(l)
    print(f, d["n_gpus"], "gpus", round(d["ms_per_step"], 3), "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["allreduce"], d["last_loss"])
except Exception as ex:
    print(f, "failed", ex); print(open(f.replace(".json", ".err")).read()[-1200:])
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2w_n1.json 2> $O/r2w_n1.err; show $O/r2w_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2w_n2.json 2> $O/r2w_n2.err; show $O/r2w_n2.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --overlap > $O/r2w_n2_overlap.json 2> $O/r2w_n2_overlap.err; show $O/r2w_n2_overlap.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --workload config3 > $O/r2w_n2_config3.json 2> $O/r2w_n2_config3.err; show $O/r2w_n2_config3.json
