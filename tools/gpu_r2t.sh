#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2t}
( timeout 1200 python -m pytest tests/test_gpu_tle.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -3 gpurun_out/${T}_pytest.log
{
for lib in "" gpurun_variants/lib_tle_nopf.so gpurun_variants/lib_tle_pfl1.so; do
  echo "== TLE lib=${lib:-default (next-voxel label prefetch to L2)}"
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle --histories 8000000 --launches 2 | cut -c1-24,100-400
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle_nosq --histories 8000000 --launches 2 | cut -c1-24,100-400
  env ${lib:+GATE_B200_LIB=$lib} GATE_B200_TLE_TABLE=1 timeout 200 python tools/profile_run.py --config c3tle --histories 4000000 --launches 2 | cut -c1-24,100-400
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-200
