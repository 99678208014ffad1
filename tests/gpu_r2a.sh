#!/bin/bash
# attention variants: correctness + timelines + A/B bench
set -u
out=gpurun_out; mkdir -p $out; tag=${1:-r2a}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "attention_on_tensor_cores or z_recursion or mdm_single or cfg_text" -s > $out/${tag}_attn_test.log 2>&1; tail -6 $out/${tag}_attn_test.log
timeout 120 python tests/attn_timeline.py > $out/${tag}_attn_timeline.log 2>&1; cat $out/${tag}_attn_timeline.log | tail -9
timeout 120 python tests/trace_probe.py > $out/${tag}_trace.log 2>&1; grep -E "^gemm_tc|^sequence|span" $out/${tag}_trace.log | cut -c1-400
ST_NO_OTHER_CONFIGS=1 ST_NO_TWO_IN_FLIGHT=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/${tag}_bench.log 2>&1; tail -1 $out/${tag}_bench.log | head -c 300; echo
