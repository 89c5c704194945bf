#!/bin/bash
# r02b (2 GPUs): gather v2 correctness (all variants), 1- and 2-GPU bench lines with the new timed train.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=${NG:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 -k "not multigpu" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
CARLB_MGPU_TIMEOUT=500 timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py > gpurun_out/mgpu_worker.log 2>&1; echo "mgpu worker exit $?"; grep -E "MGPU_OK|transport|Error|error|mismatch" gpurun_out/mgpu_worker.log | sort | uniq -c | head -30; tail -5 gpurun_out/mgpu_worker.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?"; tail -3 gpurun_out/bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench n$NG exit $?"; tail -3 gpurun_out/bench_n$NG.err
CARLB_GATHER_SYMMETRIC=ipc timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 --no-ant --no-f64 > gpurun_out/bench_n${NG}_ipc.json 2> gpurun_out/bench_n${NG}_ipc.err; echo "bench n$NG ipc exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --steps 20 --warmup 5 --no-ant --no-f64 --gather nccl > gpurun_out/bench_n${NG}_nccl.json 2> gpurun_out/bench_n${NG}_nccl.err; echo "bench n$NG nccl exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value %.4e pass_ms %.4f e2e %.4e step_api %.2f us'%(d['value'],d['config']['pass_ms_median'],d['e2e']['value'],d['step_api']['us_per_launch']), d['config']['collective'][:90], d.get('gather_check',{}).get('equal'))
    for k in ('ant_8192','config5_halfcheetah_hopper','value_f64'):
        if k in d: print('   ',k,'%.4e'%d[k]['value'], d[k].get('step_api',{}).get('us_per_launch'), d[k].get('us_per_step'))
PY
