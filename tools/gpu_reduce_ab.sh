#!/bin/bash
# A/B of the two split-reduction kernels of the row-of-taps weight gradient (run under gpurun)
mkdir -p gpurun_out
O=gpurun_out
LBC_W3_REDUCE=1 timeout 300 python -m pytest tests/test_ops.py tests/test_net_parity.py -x -q -m gpu -k "tcgen05 or bf16" > $O/test_reduce_b.log 2>&1
echo "tests (variant B) exit $?" >> $O/test_reduce_b.log
LBC_W3_REDUCE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reduce_a.json 2> $O/bench_reduce_a.err
LBC_W3_REDUCE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reduce_b.json 2> $O/bench_reduce_b.err
tail -3 $O/test_reduce_b.log; cut -c1-240 $O/bench_reduce_a.json; cut -c1-240 $O/bench_reduce_b.json
