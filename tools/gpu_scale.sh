#!/bin/bash
# Scaling lines of bench.py at N in NLIST (capped at NG) with the driver's command; prints per-N efficiency.
set +e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NG=${NG:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in ${NLIST:-1 2 4 8}; do
  if [ $n -gt $NG ]; then break; fi
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "bench n$n exit $?"; tail -2 gpurun_out/scale_n$n.err | cut -c1-300
done
python - <<'PY'
import json,glob,re
v1=None
for f in sorted(glob.glob('gpurun_out/scale_n*.json'), key=lambda s:int(re.findall(r'n(\d+)',s)[0])):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'unparsable',e); continue
    n=d['n_gpus']
    if n==1: v1=d
    def eff(a,b): return a/(n*b) if b else float('nan')
    print(f"N={n} value {d['value']:.4e} eff {eff(d['value'],v1 and v1['value']):.3f} pass_ms {d['config']['pass_ms_median']:.4f} | step_api {d['step_api']['us_per_launch']:.2f} us | e2e {d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']*1e3:.1f} us) | gather_check {d.get('gather_check',{}).get('equal')}")
    a=d.get('ant_8192'); c=d.get('config5_halfcheetah_hopper')
    if a: print(f"     ant {a['value']:.4e} (eff {eff(a['value'],v1 and v1['ant_8192']['value']):.3f}) fma {a['value_fma']:.4e} step {a['step_api']['us_per_launch']:.1f} us | config5 {c['value']:.4e} ({c['us_per_step']:.1f} us/step)")
PY
