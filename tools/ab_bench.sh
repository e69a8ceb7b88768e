#!/bin/bash
# A/B timing of experimental builds: tools/ab_bench.sh name[:kernel] ...  (libs in sitrack_b200/variants/lib_<name>.so)
for spec in "$@"; do
  n=${spec%%:*}; k=${spec##*:}; [ "$k" = "$spec" ] && k=tuned
  if [ "$n" = "base" ]; then unset SITRACK_B200_LIB; else export SITRACK_B200_LIB=$PWD/sitrack_b200/variants/lib_$n.so; fi
  timeout 100 python bench.py --steps 100 --warmup 5 --kernel $k --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$spec', d['roofline']['us_per_launch'], d['roofline']['frac'])"
done
