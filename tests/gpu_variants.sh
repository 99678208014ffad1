#!/bin/bash
# A/B of differently built libraries inside one gpurun call (same box): bash tests/gpu_variants.sh <tag> V1 V2 ...
# (syntalker_b200/variants/<V>.so are copied over the in-tree library one after the other; the last one stays)
tag=$1; shift; out=gpurun_out; mkdir -p $out
for rep in 1 2; do for v in "$@"; do
  cp syntalker_b200/variants/$v.so syntalker_b200/libsyntalker_b200.so
  timeout 300 python bench.py --steps 8 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v value %.0f ms %.3f e2e %.0f gemm_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step']))"
done; done | tee $out/${tag}_variants.log
