#!/bin/bash
# round 2, GPU call F: ncu --set full (with source) of the kernels furthest below their roof
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'conv3x3_c64|wgrad_gemm_kernel|stem_conv|stem_wgrad|bn_relu_maxpool|maxpool_relu_bwd|head_s4|head_dh4|k_head_dlogits|bn_apply_kernel' \
    --launch-skip 160 --launch-count 40 -o $O/r2f_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/r2f_full.log 2>&1
ncu -i $O/r2f_full.ncu-rep --page raw --csv > $O/r2f_full.csv 2>/dev/null
ncu -i $O/r2f_full.ncu-rep --page source --csv --kernel-name regex:conv3x3_c64 --launch-count 1 > $O/r2f_src_c64.csv 2>/dev/null
ls -la $O/r2f_full.ncu-rep
python tools/ncu_table.py full $O/r2f_full.csv | cut -c1-200
