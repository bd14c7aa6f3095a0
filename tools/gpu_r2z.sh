#!/bin/bash
# round 2, final measurements of the adopted kernels (r2z: before the TLE record variant; r2y: with it)
mkdir -p gpurun_out
T=${1:-r2z}
( timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log )
tail -4 gpurun_out/${T}_pytest.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 3000 > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_run.py 3000 > gpurun_out/${T}_sanitizer_initcheck.log 2>&1; tail -2 gpurun_out/${T}_sanitizer_initcheck.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; cut -c1-300 gpurun_out/${T}_bench_c3.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm_c3.json 2> gpurun_out/${T}_bench_reference_arm_c3.err; cut -c1-200 gpurun_out/${T}_bench_reference_arm_c3.json
for cfg in c2 c1 c4 c3_hit_random; do
  timeout 300 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/${T}_bench_${cfg}.json 2> gpurun_out/${T}_bench_${cfg}.err; cut -c1-160 gpurun_out/${T}_bench_${cfg}.json
done
timeout 300 python bench.py --config c3tle --histories 8000000 --no-cpu-baseline > gpurun_out/${T}_bench_c3tle.json 2> gpurun_out/${T}_bench_c3tle.err; cut -c1-160 gpurun_out/${T}_bench_c3tle.json
timeout 300 python bench.py --config c3tle_nosq --histories 8000000 --no-cpu-baseline > gpurun_out/${T}_bench_c3tle_nosq.json 2> gpurun_out/${T}_bench_c3tle_nosq.err; cut -c1-160 gpurun_out/${T}_bench_c3tle_nosq.json
timeout 300 python bench.py --config c2 --steps 4 --total-histories 1e10 --no-cpu-baseline --no-script-level > gpurun_out/${T}_bench_c5_c2_1gpu.json 2> gpurun_out/${T}_bench_c5_c2_1gpu.err; cut -c1-160 gpurun_out/${T}_bench_c5_c2_1gpu.json
timeout 300 python bench.py --config c5ct --steps 4 --total-histories 1e10 --no-cpu-baseline --no-script-level > gpurun_out/${T}_bench_c5_ct_1gpu.json 2> gpurun_out/${T}_bench_c5_ct_1gpu.err; cut -c1-160 gpurun_out/${T}_bench_c5_ct_1gpu.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-script-level --histories 200000000 > gpurun_out/${T}_ncu_bench.log 2>&1
for cfg in c3 c3_hit_random c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
      -o gpurun_out/${T}_transport_${cfg} -f python tools/profile_run.py --config $cfg --histories 100000000 --launches 1 --warmup 1 > gpurun_out/${T}_ncu_${cfg}.log 2>&1
  tail -1 gpurun_out/${T}_ncu_${cfg}.log
  # summarised on the box: the four reports together exceed what gpurun brings back
  python tools/ncu_summary.py gpurun_out/${T}_transport_${cfg}.ncu-rep gpurun_out/${T}_transport_${cfg}_ncu.md --histories 1e8 \
      --title "${T} ${cfg} (kernel of the bench line, 1e8 histories)" --traffic-json gpurun_out/${T}_traffic.json --key ${cfg} > /dev/null
  [ "$cfg" = c3 ] || rm -f gpurun_out/${T}_transport_${cfg}.ncu-rep
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transport_kernel -s 1 -c 1 \
    -o gpurun_out/${T}_transport_c3tle -f python tools/profile_run.py --config c3tle --histories 4000000 --launches 1 --warmup 1 > gpurun_out/${T}_ncu_c3tle.log 2>&1
tail -1 gpurun_out/${T}_ncu_c3tle.log
python tools/ncu_summary.py gpurun_out/${T}_transport_c3tle.ncu-rep gpurun_out/${T}_transport_c3tle_ncu.md --histories 4e6 \
    --title "${T} c3tle (TLE with squares, record variant, 4e6 histories)" --traffic-json gpurun_out/${T}_traffic.json --key c3tle > /dev/null
rm -f gpurun_out/${T}_transport_c3tle.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt; nproc >> gpurun_out/${T}_smi.txt
