#!/bin/bash
# A/B of a st_debug_probe experiment bit inside one gpurun call (same box): bash tests/gpu_ab.sh <tag> <bits>
tag=$1; bits=$2; out=gpurun_out; mkdir -p $out
for rep in 1 2; do for p in 0 $bits; do
  ST_PROBE=$p timeout 300 python bench.py --steps 8 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('ST_PROBE=$p value %.0f ms %.3f e2e %.0f gemm_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step']))"
done; done | tee $out/${tag}_ab.log
