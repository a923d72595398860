#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_hpr.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpr_select -s 2 -c 1 -o gpurun_out/prof_hpr python tools/run_synth_once.py > gpurun_out/prof_hpr.log 2>&1
tail -2 gpurun_out/prof_hpr.log
