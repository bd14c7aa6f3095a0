#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2s}
( timeout 1200 python -m pytest tests/test_gpu_tle.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -3 gpurun_out/${T}_pytest.log
{
echo "== record variant, ray-vs-region candidate test"
timeout 200 python tools/profile_run.py --config c3tle --histories 8000000 --launches 2 | cut -c1-24,100-400
timeout 200 python tools/profile_run.py --config c3tle_nosq --histories 8000000 --launches 2 | cut -c1-24,100-400
timeout 200 python tools/profile_run.py --config c3 --histories 100000000 --launches 2 | cut -c1-24,100-400
timeout 200 python tools/profile_run.py --config c3_hit_random --histories 100000000 --launches 2 | cut -c1-24,100-400
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o gpurun_out/${T}_transport_c3tle -f python tools/profile_run.py --config c3tle --histories 4000000 --launches 1 --warmup 1 > gpurun_out/${T}_ncu_c3tle.log 2>&1
tail -1 gpurun_out/${T}_ncu_c3tle.log
