#!/bin/bash
# Per-kernel `ncu --set full` captures of the training step on one B200 (run under gpurun; CSV exports only so that
# gpurun_out/ stays under the copy-back limit).  Usage: bash tools/gpu_profile_step.sh [kernel-regex ...]
# Default list = every kernel family of the step that is not already covered by profiles/r1_ncu_full_*.csv.
mkdir -p gpurun_out
O=gpurun_out
KERNELS=("$@")
if [ ${#KERNELS[@]} -eq 0 ]; then
  KERNELS=(maxpool_relu_bwd bn_relu_maxpool stem_conv stem_wgrad stem_pad4 k_pack_all head_s_kernel k_head_dlogits
           head_dh_kernel head_logits_kernel bn_apply_kernel ew_kernel col_finalize conv3x3_c64 "wgrad_gemm_kernel<64" k_adam)
fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 900 --csv \
    --log-file $O/launches_step.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_step.log 2>&1
for k in "${KERNELS[@]}"; do
  tag=$(echo "$k" | tr -c 'A-Za-z0-9_' '_')
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 2 --launch-count 2 \
      -o $O/ncu_$tag -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_$tag.log 2>&1
  ncu -i $O/ncu_$tag.ncu-rep --page raw --csv > $O/ncu_$tag.csv 2>/dev/null
  ncu -i $O/ncu_$tag.ncu-rep --page source --csv > $O/ncu_${tag}_source.csv 2>/dev/null
  rm -f $O/ncu_$tag.ncu-rep
done
python tools/ncu_table.py launches $O/launches_step.csv | head -40
for f in $O/ncu_*.csv; do case $f in *_source.csv) ;; *) python tools/ncu_table.py full $f | tail -n +2;; esac; done
du -sh $O
