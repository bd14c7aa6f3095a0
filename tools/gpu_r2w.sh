#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2w}
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "energy_source_types or bit_exact or box_regions" 2>&1 | tail -2
{
for rep in 1 2; do
for lib in "" gpurun_variants/lib_prev_abi4.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c4 c2 c1 c3_hit_random; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-190
