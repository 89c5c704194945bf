#!/bin/bash
# Quick GPU visit: parity tests + smoke only.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q $1 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
