#!/bin/bash
# multi-GPU call: N from the first argument.  2-GPU engine tests, bench weak (driver form) and strong (config 5)
N=${1:-2}
T=${2:-r2m}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  ( timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/${T}_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_2gpu.log )
  tail -5 gpurun_out/${T}_pytest_2gpu.log
fi
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" \
      > gpurun_out/${T}_bench_${name}_${N}gpu.json 2> gpurun_out/${T}_bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_${name}_${N}gpu.json"))
    print("${name}", "N=$N", "value %.4g e2e %.4g scaling %s reduce_check %s ms/step %.1f" % (d["value"], d["e2e"]["value"], d["scaling"], d.get("reduce_check"), d["ms_per_step"]), d["e2e"]["ms_per_step_by_section_rank0"])
except Exception as e:
    print("${name} failed", e)
    print(open("gpurun_out/${T}_bench_${name}_${N}gpu.err").read()[-1500:])
PY
}
run c3 --steps 5 --warmup 3 --no-cpu-baseline
run c5_c2 --config c2 --steps 4 --warmup 3 --total-histories 1e10 --no-cpu-baseline
run c5_ct --config c5ct --steps 4 --warmup 3 --total-histories 1e10 --no-cpu-baseline
