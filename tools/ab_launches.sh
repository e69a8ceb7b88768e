#!/bin/bash
# per-kernel durations of the step for A/B builds: [KERNEL=n] tools/ab_launches.sh name ...  (libs in sitrack_b200/variants/lib_<name>.so; "base" = the in-tree build)
for n in "$@"; do
  if [ "$n" = "base" ]; then unset SITRACK_B200_LIB; else export SITRACK_B200_LIB=$PWD/sitrack_b200/variants/lib_$n.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_advect|k_walk" --launch-skip 8 --launch-count 8 --csv --log-file /tmp/l_$n.csv python bench.py --kernel ${KERNEL:-tuned} --no-shuffle --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - "$n" /tmp/l_$n.csv <<'PY'
import csv,sys,collections
d=collections.defaultdict(list)
for r in csv.reader(open(sys.argv[2])):
    if len(r)>5 and r[-3]=="gpu__time_duration.sum": d[r[4].split("<")[0].replace("void ","")].append(float(r[-1].replace(",",""))/1e3)
print(sys.argv[1], {k:"%.1f"%(sum(v)/len(v)) for k,v in d.items()}, "sum %.1f"%sum(sum(v)/len(v) for v in d.values()))
PY
done
