#!/bin/bash
mkdir -p gpurun_out
T=r2h
( timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -5 gpurun_out/${T}_pytest.log
{
for lib in gpurun_variants/lib_base.so "" gpurun_variants/lib_t1024.so gpurun_variants/lib_t768.so gpurun_variants/lib_lazy.so; do
  echo "== lib=${lib:-default}"
  for c in c3 c2 c1 c4 c3_hit_random; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 100000000 --launches 2 | cut -c1-24,100-400; done
done
for lib in gpurun_variants/lib_base.so "" gpurun_variants/lib_tlecg.so; do
  echo "== TLE lib=${lib:-default}"
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle --histories 2000000 --launches 2 | cut -c1-24,100-400
  env ${lib:+GATE_B200_LIB=$lib} timeout 200 python tools/profile_run.py --config c3tle_nosq --histories 4000000 --launches 2 | cut -c1-24,100-400
done
} > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log | cut -c1-170
timeout 600 python bench.py > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; cut -c1-400 gpurun_out/${T}_bench_c3.json; tail -3 gpurun_out/${T}_bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 -o gpurun_out/${T}_transport_c3 -f python tools/profile_run.py --config c3 --histories 100000000 --launches 1 --warmup 1 > gpurun_out/${T}_ncu_c3.log 2>&1
