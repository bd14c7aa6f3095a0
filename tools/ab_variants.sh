#!/bin/bash
# A/B timing of alternative builds of the C-ABI library (development aid):
#   tools/ab_variants.sh "<lib1> <lib2> ..." "<configs>" [histories]
libs=${1:-""}; cfgs=${2:-"c3 c1"}; n=${3:-20000000}
for lib in "" $libs; do
  echo "== lib=${lib:-default}"
  for c in $cfgs; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories $n --launches 2 | cut -c1-24,100-400; done
done
