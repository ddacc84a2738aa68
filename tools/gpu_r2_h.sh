#!/bin/bash
# round 2, GPU call H (2 GPUs): full suite; data-parallel bench with the bucketed overlapped all-reduce vs the single one
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > $O/r2h_test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/r2h_test_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2h_n1.json 2> $O/r2h_n1.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2h_n2_overlap.json 2> $O/r2h_n2_overlap.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-overlap > $O/r2h_n2_single.json 2> $O/r2h_n2_single.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload config3 > $O/r2h_n2_config4.json 2> $O/r2h_n2_config4.err
NCCL_DEBUG=INFO timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 1 > /dev/null 2> $O/r2h_nccl_info.log
grep -h "passed\|failed" $O/r2h_test_gpu.log | tail -2
grep -h "^FAILED\|^ERROR" $O/r2h_test_gpu.log | cut -c1-200 | head -20
python - <<PY
import json
for f in ("n1", "n2_overlap", "n2_single", "n2_config4"):
    try:
        d = json.load(open("$O/r2h_%s.json" % f)); print(f, d["n_gpus"], round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), d["config"].get("allreduce"), d["config"].get("cpu_affinity"))
    except Exception as ex:
        print(f, "failed", ex); print(open("$O/r2h_%s.err" % f).read()[-1200:])
PY
grep -h "NVLS\|Channel\|via" $O/r2h_nccl_info.log | head -8 | cut -c1-200
