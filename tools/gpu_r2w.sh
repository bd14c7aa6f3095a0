#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2w}
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tle.py tests/test_fullsize_parity.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -3 gpurun_out/${T}_pytest.log
{
for lib in "" gpurun_variants/lib_nochain.so gpurun_variants/lib_chain8.so gpurun_variants/lib_chain24.so gpurun_variants/lib_chain3.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c2 c1 c4; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-190
