#!/bin/bash
# A/B timing of experimental builds: [STEPS=n] tools/ab_bench.sh name[:kernel] ...  (libs in sitrack_b200/variants/lib_<name>.so)
STEPS=${STEPS:-100}
for spec in "$@"; do
  n=${spec%%:*}; k=${spec##*:}; [ "$k" = "$spec" ] && k=tuned
  if [ "$n" = "base" ]; then unset SITRACK_B200_LIB; else export SITRACK_B200_LIB=$PWD/sitrack_b200/variants/lib_$n.so; fi
  timeout 200 python bench.py --steps $STEPS --warmup 5 --kernel $k --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$spec', d['roofline']['us_per_launch'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
