#!/bin/bash
# r02d (2 GPUs): where do the 8 us per pipelined launch go? A/B over CARLB_GATHER_DEBUG bits (measurement only).
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=${NG:-2}
run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $NG --steps 20 --warmup 5 --fused-only 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', 'value %.4e kernel_us %.2f'%(d['value'], d['kernel_ms_avg']*1e3))"
}
run baseline_mc X=1
run baseline_ipc CARLB_GATHER_SYMMETRIC=ipc
run no_cta_fence CARLB_GATHER_DEBUG=1
run no_end_wait CARLB_GATHER_DEBUG=2
run no_stores CARLB_GATHER_DEBUG=4
run no_sys_fence CARLB_GATHER_DEBUG=8
run no_fence_no_wait CARLB_GATHER_DEBUG=11
run nothing CARLB_GATHER_DEBUG=15
run ipc_no_cta_fence CARLB_GATHER_SYMMETRIC=ipc CARLB_GATHER_DEBUG=1
run ipc_no_stores CARLB_GATHER_SYMMETRIC=ipc CARLB_GATHER_DEBUG=4
timeout 200 python bench.py --steps 20 --warmup 5 --fused-only | tail -1 | cut -c1-200
# e2e: polling vs stream sync, and the split-batch async API
timeout 300 python tools/e2e_async_probe.py > gpurun_out/e2e_async_probe.json 2> gpurun_out/e2e_async_probe.err; echo "e2e probe exit $?"; cat gpurun_out/e2e_async_probe.json; tail -3 gpurun_out/e2e_async_probe.err
timeout 300 python -m pytest tests/test_env_api_gpu.py -m gpu -q -x 2>&1 | tail -3
