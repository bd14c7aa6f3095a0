#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2w}
{
for lib in "" gpurun_variants/lib_tle_s8.so gpurun_variants/lib_tle_s32.so gpurun_variants/lib_tle_s64.so; do
  echo "== TLE chords per call: lib=${lib:-default (16)}"
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle --histories 8000000 --launches 2 | cut -c1-24,100-400
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle_nosq --histories 8000000 --launches 2 | cut -c1-24,100-400
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-190
