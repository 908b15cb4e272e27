#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/sanitize_r02.py 2>&1 | tail -8
timeout 120 python tools/sanitize_r02.py 2>&1 | tail -8
timeout 600 compute-sanitizer --tool initcheck --print-limit 8 python tools/sanitize_r02.py > gpurun_out/sanitize_r02_initcheck.log 2>&1; grep -v "^=========     \|^$" gpurun_out/sanitize_r02_initcheck.log | head -40 | cut -c1-200
