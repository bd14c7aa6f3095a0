#!/bin/bash
# round 2, call e: shell crossing inside the Woodcock call: parity, A/B against the base library on one box, ncu
mkdir -p gpurun_out
T=r2e
( timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -12 gpurun_out/${T}_pytest.log
{
for lib in gpurun_variants/lib_base.so "" gpurun_variants/lib_lazy.so gpurun_variants/lib_t1024.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c2 c1 c4 c3_hit_random; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
for cfg in c3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
      -o gpurun_out/${T}_transport_${cfg} -f python tools/profile_run.py --config $cfg --histories 100000000 --launches 1 --warmup 1 \
      > gpurun_out/${T}_ncu_${cfg}.log 2>&1
  tail -1 gpurun_out/${T}_ncu_${cfg}.log
done
