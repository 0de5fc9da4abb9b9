#!/bin/bash
# A/B of the compiled-out kernel options on a GPU box (build them first, here: make -C slam-rgbd_b200 next-variants):
#   gpurun --timeout 1500 -- 'bash tools/ab_variants.sh'
# For the default library and every variant: the GPU parity tests (bit-exactness against the CPU oracle is the
# gate -- a variant that fails them is not measured), then a short bench.py run.  Results: gpurun_out/ab_<name>.*
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=slam-rgbd_b200/lib/variants
[ -x tools/bin/membench ] && tools/bin/membench > gpurun_out/membench.txt 2>&1
# the opt-in GPU leg of the mq: mode (first run)
YOUTH_TEST_MQ_MODE=1 timeout 300 python -m pytest tests/test_live_pipeline.py -x -q -m gpu > gpurun_out/mq_mode_test.log 2>&1; echo "mq mode test rc=$?"
python tools/variant_probe.py > gpurun_out/variant_probe.jsonl 2> gpurun_out/variant_probe.err; cat gpurun_out/variant_probe.jsonl | cut -c1-330
for name in default s8d2mb5 s8d3mb4 s8d4mb4 d3mb4 d4mb3 d4mb4xy3slim d5mb3xy3 d6mb3xy3; do
  if [ "$name" = default ]; then unset YOUTH_CUDA_LIB; else
    [ -f "$V/libyouth_cuda_$name.so" ] || { echo "$name: not built"; continue; }
    export YOUTH_CUDA_LIB="$PWD/$V/libyouth_cuda_$name.so"
  fi
  if timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -x -q -m gpu > "gpurun_out/ab_${name}_tests.log" 2>&1; then
    timeout 300 python bench.py --steps 30 --warmup 5 > "gpurun_out/ab_${name}_bench.json" 2> "gpurun_out/ab_${name}_bench.err"
    echo "$name: parity ok; $(python -c "import json,sys; d=json.loads(open('gpurun_out/ab_${name}_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['unit'], 'e2e', d['e2e']['value'], d.get('per_kernel_ms_per_step'))" 2>&1)"
  else
    echo "$name: PARITY FAILED (see gpurun_out/ab_${name}_tests.log)"; tail -5 "gpurun_out/ab_${name}_tests.log"
  fi
done
