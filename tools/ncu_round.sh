#!/bin/bash
# The round's ncu evidence on one B200 (each profiled command first runs plain, same arguments, exit code checked):
#   gpurun --timeout 1500 -- 'bash tools/ncu_round.sh r2h'
#  1. launch list of the bench command (gpu__time_duration.sum of every launch; cold-cache, serialised: shares only)
#  2. ncu --set full of one 300-frame launch group (k_ingest, k_normals, all k_icp launches, k_compose) at the three
#     geometries bench.py times: configs[1] (640x480, 1 sequence), configs[3] shard (8 sequences), configs[4] (1280x960, 4 levels)
set -u
cd "$(dirname "$0")/.."
tag=$1
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-extra --no-parity"
$BENCH > "gpurun_out/${tag}_bench_plain.json" 2> "gpurun_out/${tag}_bench_plain.err" &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "gpurun_out/${tag}_launches_bench_py.csv" $BENCH > "gpurun_out/${tag}_bench_ncu.log" 2>&1
echo "launch list rc=$?"
run_full() { # name, profile_step args...
  local name=$1; shift
  local CMD="python tools/profile_step.py $*"
  $CMD > "gpurun_out/${tag}_${name}_plain.log" 2>&1 &&
  ncu --set full --clock-control none -k regex:'k_ingest|k_normals|k_icp|k_compose' -c 40 -f -o "/tmp/${tag}_${name}" $CMD > "gpurun_out/${tag}_${name}_ncu.log" 2>&1
  echo "$name rc=$?"
  # gpurun_out/ travels back (64 MiB at most): keep the raw metric table of every captured launch, not the report
  ncu -i "/tmp/${tag}_${name}.ncu-rep" --page raw --csv > "gpurun_out/${tag}_${name}_raw.csv" 2> /dev/null
  rm -f "/tmp/${tag}_${name}.ncu-rep"
}
run_full main --batch 300 --groups 1 --ppt 128
# one level-0 k_icp launch and k_ingest of the main geometry WITH source correlation (per-instruction stall reasons)
CMD="python tools/profile_step.py --batch 300 --groups 1 --ppt 128"
ncu --set full --clock-control none --import-source on -k regex:k_icp -s 9 -c 1 -f -o "gpurun_out/${tag}_icp_L0" $CMD > /dev/null 2>&1; echo "icp_L0 source rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_ingest -c 1 -f -o "gpurun_out/${tag}_ingest" $CMD > /dev/null 2>&1; echo "ingest source rc=$?"
run_full s8 --batch 300 --groups 1 --ppt 128 --streams 8
run_full hires --batch 300 --groups 1 --ppt 128 --width 1280 --height 960 --levels 4
