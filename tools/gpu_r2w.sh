#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2w}
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tle.py tests/test_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -3 gpurun_out/${T}_pytest.log
{
for rep in 1 2; do
for lib in "" gpurun_variants/lib_noclass.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c2 c1 c4 c3tle; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories $([ $c = c3tle ] && echo 8000000 || echo 100000000) --launches 2 | cut -c1-24,100-400; done
done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-190
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o gpurun_out/${T}_transport_c3 -f python tools/profile_run.py --config c3 --histories 100000000 --launches 1 --warmup 1 > gpurun_out/${T}_ncu_c3.log 2>&1
tail -1 gpurun_out/${T}_ncu_c3.log
