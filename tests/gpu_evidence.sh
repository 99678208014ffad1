#!/bin/bash
# One gpurun call that collects the round's evidence set (every call costs about a minute of box time before the command
# starts, so the pieces are batched):   gpurun --timeout 2400 -- 'bash tests/gpu_evidence.sh r2'
# Outputs go to gpurun_out/<tag>_*; gpurun brings back at most 64 MiB, so every `ncu --set full` report is reduced on the box to
# its raw-page CSV + the compact table of tests/ncu_extract.py, and only the trunk-kernel report itself is kept.
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout 500 python bench.py --steps 8 --warmup 3 > $out/${tag}_bench.log 2>&1; tail -1 $out/${tag}_bench.log > $out/${tag}_bench.json; head -c 300 $out/${tag}_bench.json; echo
timeout 120 python tests/trace_probe.py > $out/${tag}_trace.log 2>&1; grep "^sequence" $out/${tag}_trace.log | cut -c1-300
timeout 120 python tests/attn_timeline.py > $out/${tag}_attn_timeline.log 2>&1; tail -9 $out/${tag}_attn_timeline.log
timeout 120 python tests/trace_pipeline.py > $out/${tag}_trace_pipeline.log 2>&1; grep -E "^--|total" $out/${tag}_trace_pipeline.log | head
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $out/${tag}_launches.csv python tests/ncu_target.py tc 32 > $out/${tag}_ncu_list.log 2>&1
python tests/agg_launches.py $out/${tag}_launches.csv > $out/${tag}_launches_summary.txt; head -8 $out/${tag}_launches_summary.txt
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f"
cap() {   # name, kernel regex, extra ncu args, target section, keep-report flag
  timeout 300 $NCU -k "regex:$2" $3 -o $out/${tag}_$1 python tests/ncu_target2.py $4 > $out/${tag}_ncu_$1.log 2>&1
  ncu -i $out/${tag}_$1.ncu-rep --page raw --csv > $out/${tag}_$1_raw.csv 2>/dev/null
  python tests/ncu_extract.py $out/${tag}_$1_raw.csv $out/${tag}_ncu_$1.csv
  gzip -f $out/${tag}_$1_raw.csv
  [ "$5" = keep ] || rm -f $out/${tag}_$1.ncu-rep
}
cap step_fast gemm_tc_fast_kernel "-c 14" step keep
cap step_tokens tokens_step_kernel "-c 2" step drop
cap decode_misc "vq_select_kernel|pose330_kernel|trans_kernel|gemm_tc_kernel" "-c 10" decode drop
cap decode_conv gemm_tc_fast_kernel "-s 2 -c 6" decode drop
cap encode_simt "gemm_simt_kernel|wav_first_kernel" "-c 6" encode drop
cap encode_tc gemm_tc_kernel "-c 8" encode drop
timeout 170 compute-sanitizer --tool memcheck python tests/sanitize_target.py > $out/${tag}_san_mem.log 2>&1; tail -1 $out/${tag}_san_mem.log
du -sh $out
