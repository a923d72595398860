#!/bin/bash
# HPR group solver with explicit warp ordering: tests, racecheck, step time — each under its own timeout
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_synthesis.py -x -q 2>&1 | tail -2
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_synthesis.py -k 'test_hidden_point_removal_matches_qhull or test_hpr_duplicates_and_padding_draws' -x -q -p no:cacheprovider > gpurun_out/sanitize_racecheck_hpr.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_racecheck_hpr.log; grep -c "Race reported" gpurun_out/sanitize_racecheck_hpr.log
timeout 60 python tools/ab_pipeline.py 1 2>&1 | tail -1
