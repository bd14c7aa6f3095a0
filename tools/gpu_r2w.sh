#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2w}
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bit_exact or uint16" 2>&1 | tail -2
{
for rep in 1 2; do
for lib in "" gpurun_variants/lib_alloc.so; do
  echo "== lib=${lib:-default (labels loaded with L1::no_allocate)}"
  for c in c3 c2 c1 c4; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-190
