#!/bin/bash
# Runs every GEMM mode in its own process (a trap poisons the context) with a per-case timeout.
cd "$(dirname "$0")/.."
out=gpurun_out/gemm_diag.log
mkdir -p gpurun_out; : > $out
run() { timeout 120 python tools/gemm_diag.py "$@" >> $out 2>&1; echo "exit=$? args=$*" >> $out; }
run 0 0 128 1 128 128 64
run 0 0 128 1 256 256 256
run 0 0 256 1 384 512 768
run 0 0 64 1 200 72 136
run 0 0 0 1 1000 2304 768
run 0 1 128 1 128 128 64
run 0 1 256 1 384 512 768
run 0 1 0 1 1000 768 2304
run 1 1 128 1 128 128 64
run 1 1 256 1 384 512 768
run 1 1 0 0 2304 768 5000
run 1 0 128 1 256 256 256
run 0 0 128 4 256 256 1024
cat $out
