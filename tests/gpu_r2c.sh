#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=${1:-r2c}
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 -s > $out/${tag}_pytest.log 2>&1; tail -40 $out/${tag}_pytest.log | cut -c1-300
ST_NO_TWO_IN_FLIGHT=1 timeout 300 python bench.py --steps 8 --warmup 3 > $out/${tag}_bench.log 2>&1; tail -1 $out/${tag}_bench.log | head -c 300; echo
