#!/bin/bash
mkdir -p gpurun_out
T=r2k
{
for lib in "" gpurun_variants/lib_pp1.so gpurun_variants/lib_pp2.so gpurun_variants/lib_pp4.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c2 c1 c4 c3_hit_random; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-170
