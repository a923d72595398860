#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/b128_parity.json
( time timeout 900 python -m pytest tests/test_gpu_model_b128.py -q -s ) > gpurun_out/pytest_b128.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_b128.log
( CAAE_TEST_PRECISION=fp32 timeout 900 python -m pytest tests/test_gpu_model_b128.py -q -s -k "train and dgcnn" ) > gpurun_out/pytest_b128_fp32.log 2>&1
grep "B=128" gpurun_out/pytest_b128.log gpurun_out/pytest_b128_fp32.log
tail -5 gpurun_out/pytest_b128.log
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_model_b128.py ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
