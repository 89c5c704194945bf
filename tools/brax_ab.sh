#!/bin/bash
# A/B of libcarlb.so variants under build_variants/*_libcarlb.so (built on the dev box; they travel with gpurun):
# throughput table + dynamic instruction counts of the Ant kernels per variant. Restores the tree's own library.
set +e
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
cp carl_b200/csrc/libcarlb.so /tmp/own_libcarlb.so
for so in build_variants/*_libcarlb.so; do
  v=$(basename $so _libcarlb.so)
  cp $so carl_b200/csrc/libcarlb.so
  echo "== $v"; timeout 200 python tools/brax_ab.py 2>gpurun_out/ab_$v.err | tee gpurun_out/ab_$v.json
  for arith in strict fma; do
    timeout 120 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:brax_step_kernel -c 3 --csv \
      python tools/brax_ab.py --ncu ant $arith 2>/dev/null | grep -E "brax_step_kernel" | awk -F'","' '{print "'$v' ant '$arith'", $(NF-2), $(NF)}' | tail -2
  done
done
cp /tmp/own_libcarlb.so carl_b200/csrc/libcarlb.so
