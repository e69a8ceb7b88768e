#!/usr/bin/env python3
"""profiles/ncu_traffic.json from a tools/ncu_summary.py JSON:  update_traffic.py <summary.json> <workload> <source note>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
summ, workload, note = sys.argv[1], sys.argv[2], sys.argv[3]
k = json.load(open(summ))[0]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
t = json.load(open(path))
t[workload] = {"dram_bytes_per_launch": int(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]),
               "kernel": k["kernel"].replace("void ", "").split("(")[0], "source": note}
t["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the default step kernel from ncu --set full; "
                 "read by bench.py for roofline.traffic")
json.dump(t, open(path, "w"), indent=1)
print(t[workload])
