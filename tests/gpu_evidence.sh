#!/bin/bash
# One gpurun call that collects the round's evidence set (every call costs about a minute of box time before the command
# starts, so the pieces are batched):   gpurun --timeout 1500 -- 'bash tests/gpu_evidence.sh r2'
# Outputs go to gpurun_out/<tag>_*; copy what should be judged into profiles/ (tests/agg_launches.py, tests/ncu_extract.py).
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout 400 python bench.py --steps 8 --warmup 3 > $out/${tag}_bench.log 2>&1; tail -1 $out/${tag}_bench.log > $out/${tag}_bench.json; head -c 400 $out/${tag}_bench.json; echo
timeout 120 python tests/trace_probe.py > $out/${tag}_trace.log 2>&1; grep "^sequence" $out/${tag}_trace.log | cut -c1-300
timeout 120 python tests/attn_timeline.py > $out/${tag}_attn_timeline.log 2>&1; tail -9 $out/${tag}_attn_timeline.log
timeout 120 python tests/trace_pipeline.py > $out/${tag}_trace_pipeline.log 2>&1; tail -12 $out/${tag}_trace_pipeline.log
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $out/${tag}_launches.csv python tests/ncu_target.py tc 32 > $out/${tag}_ncu_list.log 2>&1
python tests/agg_launches.py $out/${tag}_launches.csv | head -12
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f"
timeout 240 $NCU -k regex:gemm_tc_fast_kernel -c 14 -o $out/${tag}_step_fast python tests/ncu_target2.py step > $out/${tag}_ncu_a.log 2>&1; tail -1 $out/${tag}_ncu_a.log
timeout 240 $NCU -k regex:tokens_step_kernel -c 2 -o $out/${tag}_step_tokens python tests/ncu_target2.py step > $out/${tag}_ncu_b.log 2>&1; tail -1 $out/${tag}_ncu_b.log
timeout 240 $NCU -k "regex:vq_select_kernel|pose330_kernel|trans_kernel|gemm_tc_kernel" -c 10 -o $out/${tag}_decode_misc python tests/ncu_target2.py decode > $out/${tag}_ncu_c.log 2>&1; tail -1 $out/${tag}_ncu_c.log
timeout 240 $NCU -k regex:gemm_tc_fast_kernel -s 2 -c 6 -o $out/${tag}_decode_conv python tests/ncu_target2.py decode > $out/${tag}_ncu_d.log 2>&1; tail -1 $out/${tag}_ncu_d.log
timeout 240 $NCU -k "regex:gemm_simt_kernel|wav_first_kernel" -c 6 -o $out/${tag}_encode_simt python tests/ncu_target2.py encode > $out/${tag}_ncu_e.log 2>&1; tail -1 $out/${tag}_ncu_e.log
timeout 240 $NCU -k regex:gemm_tc_kernel -c 8 -o $out/${tag}_encode_tc python tests/ncu_target2.py encode > $out/${tag}_ncu_f.log 2>&1; tail -1 $out/${tag}_ncu_f.log
timeout 170 compute-sanitizer --tool memcheck python tests/sanitize_target.py > $out/${tag}_san_mem.log 2>&1; tail -1 $out/${tag}_san_mem.log
ls -la $out | grep ${tag}_ | awk '{print $5, $9}'
