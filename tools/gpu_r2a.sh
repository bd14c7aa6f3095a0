#!/bin/bash
# round 2, call a: GPU test tier, ncu captures of the kernel variants the bench times, baseline bench
mkdir -p gpurun_out
T=r2a
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -5 gpurun_out/${T}_pytest.log
for cfg in c3 c3_hit_random c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
      -o gpurun_out/${T}_transport_${cfg} -f python tools/profile_run.py --config $cfg --histories 100000000 --launches 1 --warmup 1 \
      > gpurun_out/${T}_ncu_${cfg}.log 2>&1
  tail -2 gpurun_out/${T}_ncu_${cfg}.log
done
timeout 600 python bench.py > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
cat gpurun_out/${T}_bench_c3.json | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --histories 200000000 > gpurun_out/${T}_ncu_bench.log 2>&1
for cfg in c3_hit_random c2 c1 c4; do
  timeout 300 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/${T}_bench_${cfg}.json 2> gpurun_out/${T}_bench_${cfg}.err
  cut -c1-200 gpurun_out/${T}_bench_${cfg}.json
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt; nproc >> gpurun_out/${T}_smi.txt
