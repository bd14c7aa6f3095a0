#!/bin/bash
# A/B timing of builds that run 2 CTAs per SM (development aid): tools/ab_ctas.sh "<lib:ctas> ..." "<configs>" [histories]
libs=${1:-""}; cfgs=${2:-"c3 c1"}; n=${3:-50000000}
for spec in $libs; do
  lib=${spec%%:*}; ctas=${spec#*:}
  echo "== lib=$lib ctas=$ctas"
  for c in $cfgs; do env GATE_B200_LIB=$lib GATE_B200_CTAS=$ctas timeout 120 python tools/profile_run.py --config $c --histories $n --launches 2 | cut -c1-24,100-400; done
done
