#!/bin/bash
# round 2, GPU call M: full GPU suite on the tree with the s2d stem + mma.sync heads; ncu --set full of the stem / pool /
# head kernels and of a dozen conv GEMM launches (DRAM traffic of the dominant kernel for bench.py's roofline.traffic)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > $O/r2m_test.log 2>&1
echo "exit $?" >> $O/r2m_test.log
grep -h "passed\|failed" $O/r2m_test.log | tail -2
grep -h "^FAILED\|^ERROR\|^E  " $O/r2m_test.log | cut -c1-300 | head -20
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'bn_relu_maxpool|maxpool_relu_bwd|stem_pad4|stem_conv|stem_wgrad|head_.*mma' \
    --launch-skip 8 --launch-count 8 -o $O/r2m_small -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r2m_small.log 2>&1
ncu -i $O/r2m_small.ncu-rep --page raw --csv > $O/r2m_small.csv 2>/dev/null
for k in bn_relu_maxpool maxpool_relu_bwd head_dh_mma; do
  ncu -i $O/r2m_small.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > $O/r2m_src_$k.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none \
    -k regex:'conv_gemm_kernel|conv3x3_c64|wgrad3_gemm|bn_apply_kernel|bn_bwd_apply|bn_finalize' \
    --launch-skip 330 --launch-count 60 -o $O/r2m_conv -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r2m_conv.log 2>&1
ncu -i $O/r2m_conv.ncu-rep --page raw --csv > $O/r2m_conv.csv 2>/dev/null
ls -la $O/r2m_small.ncu-rep $O/r2m_conv.ncu-rep
python tools/ncu_table.py full $O/r2m_small.csv | cut -c1-220
python tools/ncu_table.py full $O/r2m_conv.csv | cut -c1-220 | head -70
