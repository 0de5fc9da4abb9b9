#!/bin/bash
# ncu --set full of one level-0 k_icp launch (300 pairs, icp_ppt 128) for the default library and the named variants:
#   gpurun --timeout 900 -- 'bash tools/ncu_variants.sh r2c s8d4mb4 d3mb4'
set -u
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p gpurun_out
CMD="python tools/profile_step.py --batch 300 --groups 1 --ppt 128"
for name in default "$@"; do
  if [ "$name" = default ]; then unset YOUTH_CUDA_LIB; else export YOUTH_CUDA_LIB="$PWD/slam-rgbd_b200/lib/variants/libyouth_cuda_$name.so"; fi
  $CMD > "gpurun_out/${tag}_${name}_plain.log" 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_icp -s 9 -c 1 -f -o "gpurun_out/${tag}_icp_${name}" $CMD > "gpurun_out/${tag}_${name}_ncu.log" 2>&1
  echo "$name rc=$?"
done
