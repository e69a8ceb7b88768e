#!/bin/bash
# A/B timing of the chained step: [STEPS=n] tools/ab_chain.sh name[:row_buffers[:extra bench flag]] ...
# (libs in sitrack_b200/variants/lib_<name>.so; "base" = the in-tree build; row_buffers 1 = rows stepped in place)
STEPS=${STEPS:-100}
for spec in "$@"; do
  IFS=: read -r n nb flag <<< "$spec"
  if [ "$n" = "base" ]; then unset SITRACK_B200_LIB; else export SITRACK_B200_LIB=$PWD/sitrack_b200/variants/lib_$n.so; fi
  timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --e2e-steps 4 --row-buffers ${nb:-2} $flag 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$spec', d['roofline']['us_per_launch'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
