#!/bin/bash
mkdir -p gpurun_out
T=r2o
{
for lib in "" gpurun_variants/lib_tle1280.so gpurun_variants/lib_tle640.so gpurun_variants/lib_tle320.so; do
  echo "== TLE lib=${lib:-default}"
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle --histories 4000000 --launches 2 | cut -c1-24,100-400
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle_nosq --histories 8000000 --launches 2 | cut -c1-24,100-400
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-170
