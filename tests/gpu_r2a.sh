#!/bin/bash
# tensor-core attention (probe 2048) correctness + timelines, A/B bench
set -u
out=gpurun_out; mkdir -p $out; tag=${1:-r2a}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "attention_on_tensor_cores or z_recursion or mdm_single or gemm_engine" -s > $out/${tag}_attn_test.log 2>&1; tail -8 $out/${tag}_attn_test.log
for p in 2048; do
  ST_PROBE=$p timeout 120 python tests/attn_timeline.py > $out/${tag}_attn_timeline_$p.log 2>&1; cat $out/${tag}_attn_timeline_$p.log | tail -10
  ST_PROBE=$p timeout 120 python tests/trace_probe.py > $out/${tag}_trace_$p.log 2>&1; grep -E "^gemm_tc|^sequence|span" $out/${tag}_trace_$p.log | cut -c1-400
done
ST_NO_TWO_IN_FLIGHT=1 timeout 300 python bench.py --steps 8 --warmup 3 > $out/${tag}_bench.log 2>&1; tail -1 $out/${tag}_bench.log | head -c 300; echo
ST_NO_TWO_IN_FLIGHT=1 ST_PROBE=2048 timeout 300 python bench.py --steps 8 --warmup 3 > $out/${tag}_bench_2048.log 2>&1; tail -1 $out/${tag}_bench_2048.log | head -c 300; echo
