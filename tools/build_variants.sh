#!/bin/bash
# Experiment: rebuild the library with extra nvcc flags (attention / up-sampling build options) on the GPU box and bench each variant.
# usage: tools/attn_variants.sh "<flags variant 1>" "<flags variant 2>" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  touch s3od_b200/csrc/attention.cuh s3od_b200/csrc/elementwise.cuh s3od_b200/csrc/kernels_misc.cu
  S3OD_NVCC_FLAGS="$flags" python -m s3od_b200.build > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  python bench.py --cpu-sample 0 --steps 8 --e2e-steps 2 --no-extras --no-gpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
a=d['kernels']['attention']
print('$flags', '| images/s', d['value'], '| e2e', d['e2e']['value'], '| attention ms/img', a['ms_per_image'], 'TF/s', a['tflops'], '| head_bandwidth ms/img', d['kernels']['head_bandwidth']['ms_per_image'], '| layernorm', d['kernels']['layernorm']['ms_per_image'], '| b1 ms', d['latency_b1_ms'], '| verified', d['verified']['batch_vs_single_bitwise'], '| sm_mhz', d['clocks']['sm_mhz'])"
done
touch s3od_b200/csrc/attention.cuh s3od_b200/csrc/elementwise.cuh s3od_b200/csrc/kernels_misc.cu
python -m s3od_b200.build > /dev/null 2>&1
