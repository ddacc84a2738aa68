#!/bin/bash
# Round-end validation on one B200 (run under gpurun): GPU suite, smoke, bench (both arms), launch list, ncu --set full
# extracts (CSV only; .ncu-rep files are deleted to stay under the copy-back limit).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/test_gpu.log 2>&1
echo "gpu suite exit $?" >> $O/test_gpu.log
timeout 400 python __graft_entry__.py smoke > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/smoke.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_final.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source off -k regex:'wgrad3_gemm|bn_bwd_apply|bn_bwd_reduce|conv3x3_c64|wgrad_gemm' \
    --launch-skip 20 --launch-count 24 -o $O/ncu_final -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_final.log 2>&1
ncu -i $O/ncu_final.ncu-rep --page raw --csv > $O/ncu_final.csv 2>/dev/null
rm -f $O/ncu_final.ncu-rep
du -sh $O
tail -3 $O/test_gpu.log; tail -4 $O/smoke.log; cut -c1-400 $O/bench.json; cut -c1-300 $O/bench_ref.json
