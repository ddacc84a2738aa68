#!/bin/bash
# A/B of the kernel variants on one B200 (run under gpurun): op parity of every variant, bench with the switches
# off / pair / pair+wgrad3, ncu (CSV only; the .ncu-rep files are deleted to stay under the copy-back limit), the
# whole GPU suite with the variants on, and a launch list of one step.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout 420 python -m pytest tests/test_ops.py -x -q -m gpu > $O/test_ops.log 2>&1
echo "ops tests exit $?" >> $O/test_ops.log
for m in 0 1 3; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --pair $m > $O/bench_v$m.json 2> $O/bench_v$m.err
done
LBC_PAIR=3 timeout 900 python -m pytest tests -x -q -m gpu > $O/test_gpu_all_v3.log 2>&1
echo "gpu suite (LBC_PAIR=3) exit $?" >> $O/test_gpu_all_v3.log
for m in 0 3; do
  timeout 420 ncu --set full --clock-control none --import-source off -k regex:'conv_gemm_kernel|wgrad' --launch-skip 8 --launch-count 14 \
      -o $O/ncu_v$m -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --pair $m > $O/ncu_v$m.log 2>&1
  ncu -i $O/ncu_v$m.ncu-rep --page raw --csv > $O/ncu_v$m.csv 2>/dev/null
  rm -f $O/ncu_v$m.ncu-rep
done
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/launches_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pair 3 > $O/launches_v3.log 2>&1
du -sh $O
tail -3 $O/test_ops.log; for m in 0 1 3; do cut -c1-330 $O/bench_v$m.json; done; tail -3 $O/test_gpu_all_v3.log
