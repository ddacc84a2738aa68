#!/bin/bash
# A/B of the kernel variants on one B200 (run under gpurun).  LBC_PAIR bits: 1 = CTA-pair conv GEMMs, 2 = row-of-taps
# weight gradient, 4 = its CTA-pair variant.  Op parity of every variant, the whole GPU suite with everything on,
# bench per mode, and a launch list of one step in the last mode.
mkdir -p gpurun_out
O=gpurun_out
MODES=${MODES:-"0 1 3 7"}
timeout 420 python -m pytest tests/test_ops.py -x -q -m gpu > $O/test_ops.log 2>&1
echo "ops tests exit $?" >> $O/test_ops.log
LBC_PAIR=7 timeout 900 python -m pytest tests -x -q -m gpu > $O/test_gpu_all_v7.log 2>&1
echo "gpu suite (LBC_PAIR=7) exit $?" >> $O/test_gpu_all_v7.log
for m in $MODES; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --pair $m > $O/bench_v$m.json 2> $O/bench_v$m.err
done
last=${MODES##* }
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/launches_v$last.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pair $last > $O/launches_v$last.log 2>&1
du -sh $O
tail -3 $O/test_ops.log; for m in $MODES; do cut -c1-330 $O/bench_v$m.json; done; tail -3 $O/test_gpu_all_v7.log
