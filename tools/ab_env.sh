#!/bin/bash
# A/B of a run-time switch: tools/ab_env.sh VAR v1 v2 ...   (STEPS=n for the step count)
STEPS=${STEPS:-100}; var=$1; shift
for v in "$@"; do
  if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
  timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', d['roofline']['us_per_launch'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['e2e']['value'])"
done
