#!/bin/bash
mkdir -p gpurun_out
T=r2g
{
for lib in "" gpurun_variants/lib_l1.so gpurun_variants/lib_split1.so gpurun_variants/lib_split2.so gpurun_variants/lib_split3.so gpurun_variants/lib_split2l1.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c2 c1 c4; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-160
# parity of the two kinds of variant
GATE_B200_LIB=gpurun_variants/lib_split2l1.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tle.py -m gpu -q -x 2>&1 | tail -3
