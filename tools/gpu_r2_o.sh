#!/bin/bash
# round 2, GPU call O: CUDA-graph inference path (lbc_net_infer): parity with the eager eval forward + latency
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_net_parity.py tests/test_records.py -q -m gpu --tb=short -x -k "inference_graph or eval" > $O/r2o_test.log 2>&1
echo "exit $?" >> $O/r2o_test.log
grep -h "passed\|failed" $O/r2o_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2o_test.log | cut -c1-300 | head -20
timeout 300 python tools/gpu_infer_probe.py > $O/r2o_infer.jsonl 2> $O/r2o_infer.err; cat $O/r2o_infer.jsonl; tail -3 $O/r2o_infer.err
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > $O/r2o_test_all.log 2>&1
grep -h "passed\|failed" $O/r2o_test_all.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2o_test_all.log | cut -c1-300 | head -20
