#!/bin/bash
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; echo "probe exit $?"; cat gpurun_out/e2e_probe.json; tail -3 gpurun_out/e2e_probe.err
