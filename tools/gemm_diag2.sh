#!/bin/bash
# CTA-pair GEMM (bn=512 forces it): every operand layout in its own process with a timeout.
cd "$(dirname "$0")/.."
out=gpurun_out/gemm_diag2.log
mkdir -p gpurun_out; : > $out
run() { timeout 120 python tools/gemm_diag.py "$@" >> $out 2>&1; echo "exit=$? args=$*" >> $out; }
run 0 0 512 1 2048 256 64
run 0 0 512 1 4096 512 768
run 0 0 512 1 2500 328 136
run 0 1 512 1 4096 768 2304
run 1 1 512 1 2304 768 4096
run 1 1 512 0 2304 768 16384
run 1 0 512 1 2048 512 256
run 0 0 512 3 2048 512 1024
cat $out
