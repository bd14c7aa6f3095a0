#!/bin/bash
# round 2: shell-start Woodcock calls in a queue of their own (GB2_SPLIT_SHELL) -- parity, then A/B
mkdir -p gpurun_out
T=${1:-r2r}
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -4 gpurun_out/${T}_pytest.log
{
for lib in "" gpurun_variants/lib_nosplit.so gpurun_variants/lib_t1024.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c3_hit_random c4 c2 c1; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
echo "== default, L2 fetch granularity 32"
for c in c3 c2; do GATE_B200_L2_FETCH=32 timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
echo "== default again"
for c in c3 c2; do timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-190
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o gpurun_out/${T}_transport_c3 -f python tools/profile_run.py --config c3 --histories 100000000 --launches 1 --warmup 1 > gpurun_out/${T}_ncu_c3.log 2>&1
tail -1 gpurun_out/${T}_ncu_c3.log
