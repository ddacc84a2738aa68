#!/bin/bash
# round 2, GPU call B: BatchNorm-backward triage, the fp32tc (split-precision tensor-core) mode: per-op, whole net, smoke
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/gpu_dbg_bn.py > $O/r2b_dbg_bn.log 2>&1
timeout 900 python -m pytest tests/test_ops.py tests/test_kernels.py -q -m gpu --tb=short -s -k "fp32tc or stem or deconv or block_entry or bn_backward" > $O/r2b_test_ops.log 2>&1
echo "exit $?" >> $O/r2b_test_ops.log
timeout 900 python -m pytest tests/test_net_parity.py -q -m gpu --tb=short -s -k "fp32tc or taps" > $O/r2b_test_net.log 2>&1
echo "exit $?" >> $O/r2b_test_net.log
timeout 600 python __graft_entry__.py smoke > $O/r2b_smoke.log 2>&1
echo "exit $?" >> $O/r2b_smoke.log
timeout 400 python bench.py --precision fp32tc --batch 64 --steps 3 --warmup 1 --no-cpu-baseline > $O/r2b_bench_tc64.json 2> $O/r2b_bench_tc64.err
cat $O/r2b_dbg_bn.log | cut -c1-400
grep -h "passed\|failed" $O/r2b_test_ops.log $O/r2b_test_net.log | tail -4
grep -h "^FAILED\|^ERROR" $O/r2b_test_ops.log $O/r2b_test_net.log | cut -c1-200 | head -40
grep -h "^fp32tc\|^taps" $O/r2b_test_ops.log $O/r2b_test_net.log | cut -c1-600 | head -60
tail -5 $O/r2b_smoke.log | cut -c1-300
cut -c1-600 $O/r2b_bench_tc64.json; tail -3 $O/r2b_bench_tc64.err | cut -c1-300
