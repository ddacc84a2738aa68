#!/bin/bash
# short round-end check on one B200 (run under gpurun): GPU suite, smoke, bench
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $O/test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/test_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
tail -3 $O/test_gpu.log; tail -4 $O/smoke.log; cut -c1-300 $O/bench.json; python -c "
import json; d=json.load(open('$O/bench.json')); print('e2e', d['e2e']['value'], d['e2e']['fp32_frames']['value'], 'frac', d['roofline']['frac'])"
