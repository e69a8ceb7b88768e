#!/bin/bash
# Round-end check on one B200: ncu capture of the default (chained) step first (its DRAM bytes go into
# profiles/ncu_traffic.json, which bench.py quotes), GPU suite, both bench arms the way the driver runs them, launch list.
O=gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-allgather"
timeout 120 $B > $O/r2c_plain.json 2> $O/r2c_plain.err || { echo "bench failed"; tail -5 $O/r2c_plain.err; exit 1; }
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_advect_warp --launch-skip 5 --launch-count 1 \
    -f -o $O/prof_r02_chain $B > $O/ncu_r02_chain.log 2>&1
python tools/ncu_summary.py $O/prof_r02_chain.ncu-rep $O/r02_chain > /dev/null 2>&1 && \
    python tools/update_traffic.py $O/r02_chain.json cfg5 "profiles/r02_chain.json (ncu --set full, chained default step, round 2)" && \
    cp profiles/ncu_traffic.json $O/ncu_traffic.json
timeout 100 python __graft_entry__.py smoke > $O/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2c_smoke.log
timeout 330 python -m pytest tests -m gpu -x -q -n 3 --durations=12 > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2c_pytest.log
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2c_bench_ref.json 2> $O/r2c_bench_ref.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/r2c_bench.json')); print(d['value'], d['roofline'], d['e2e']['value'], d['clocks'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_launches.csv $B > /dev/null 2>&1
timeout 200 python bench.py > $O/r2c_bench_default.json 2> $O/r2c_bench_default.err
python -c "import json; d=json.load(open('$O/r2c_bench_default.json')); print('default', d['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['clocks'])"
