#!/bin/bash
# one ncu capture (speed-of-light, memory, launch, occupancy, warp-state sections) of every kernel class (tools/ncu_kernels_driver.py); summarise here with tools/ncu_table.py
set -u
mkdir -p gpurun_out
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis --clock-control none --profile-from-start off -k regex:'^k_[^s]|^k_s[^t]' -f -o gpurun_out/all_kernels_${1:-r02} python tools/ncu_kernels_driver.py > gpurun_out/ncu_all.log 2>&1
tail -3 gpurun_out/ncu_all.log | cut -c1-120
# the report is ~90 MB (gpurun_out/ is capped at 64 MiB): summarise it on the box and bring back the tables only
python tools/ncu_table.py gpurun_out/all_kernels_${1:-r02}.ncu-rep > gpurun_out/ncu_all_kernels_${1:-r02}.txt 2>&1
rm -f gpurun_out/all_kernels_${1:-r02}.ncu-rep
cat gpurun_out/ncu_all_kernels_${1:-r02}.txt
