#!/bin/bash
# ncu --set full of the 4th k_mmq_tc launch of the shape bench (gate/up q4_K is the 4th shape -> launch index 3*13.. use -s)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_mmq_tc -s ${SKIP:-45} -c 1 -f -o gpurun_out/k_mmq_tc_${1:-r01} \
    python tools/prefill_bench.py 2048 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
