#!/bin/bash
for f in ${FLAGS:-0 1 3 7}; do echo "== B200_TC_FLAGS=$f"; B200_TC_FLAGS=$f timeout 100 python tools/prefill_bench.py 2048 2>&1 | grep -v "Warning\|^{" | head -6; done
