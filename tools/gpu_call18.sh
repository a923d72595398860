#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32' --launch-skip 4 -c 4 -o gpurun_out/prof_gemm python tools/gemm_once.py > gpurun_out/prof_gemm.log 2>&1
tail -3 gpurun_out/prof_gemm.log; ls -la gpurun_out/prof_gemm.ncu-rep
