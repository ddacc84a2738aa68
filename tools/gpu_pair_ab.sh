#!/bin/bash
# A/B of the CTA-pair implicit-GEMM kernels on one B200 (run under gpurun): parity tests of the variant, bench with the
# switch off / on, ncu --set full of the first forward conv GEMMs in both modes, then the whole GPU test suite.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout 420 python -m pytest tests/test_ops.py -x -q -m gpu -k "tcgen05" > $O/test_pair.log 2>&1
echo "pair tests exit $?" >> $O/test_pair.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pair 0 > $O/bench_pair0.json 2> $O/bench_pair0.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pair 1 > $O/bench_pair1.json 2> $O/bench_pair1.err
for m in 0 1; do
  timeout 420 ncu --set full --clock-control none --import-source off -k regex:conv_gemm_kernel --launch-count 26 \
      -o $O/ncu_convgemm_pair$m -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --pair $m > $O/ncu_pair$m.log 2>&1
  ncu -i $O/ncu_convgemm_pair$m.ncu-rep --page raw --csv > $O/ncu_convgemm_pair$m.csv 2>/dev/null
done
timeout 900 python -m pytest tests -x -q -m gpu > $O/test_gpu_all.log 2>&1
echo "gpu suite exit $?" >> $O/test_gpu_all.log
tail -3 $O/test_pair.log; cat $O/bench_pair0.json | cut -c1-400; cat $O/bench_pair1.json | cut -c1-400; tail -3 $O/test_gpu_all.log
