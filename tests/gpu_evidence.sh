#!/bin/bash
# One gpurun call that collects the round's evidence set (every call costs about a minute of box time before the command
# starts, so the pieces are batched):   gpurun --timeout 900 -- 'bash tests/gpu_evidence.sh r2a'
# Outputs go to gpurun_out/<tag>_*; copy what should be judged into profiles/ (tests/agg_launches.py, tests/ncu_extract.py).
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout 500 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 300 python bench.py --steps 8 --warmup 3 > $out/${tag}_bench.log 2>&1; tail -1 $out/${tag}_bench.log | head -c 300; echo
timeout 120 python tests/trace_probe.py > $out/${tag}_trace.log 2>&1; grep "^sequence" $out/${tag}_trace.log | cut -c1-300
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $out/${tag}_launches.csv python tests/ncu_target.py tc 32 > $out/${tag}_ncu_list.log 2>&1
python tests/agg_launches.py $out/${tag}_launches.csv | head -8
timeout 240 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tc_fast_kernel -s 40 -c 8 -f \
  -o $out/${tag}_fast python tests/ncu_target.py tc 32 ddim5 > $out/${tag}_ncu_full.log 2>&1; tail -1 $out/${tag}_ncu_full.log
timeout 170 compute-sanitizer --tool memcheck python tests/sanitize_target.py > $out/${tag}_san_mem.log 2>&1; tail -1 $out/${tag}_san_mem.log
