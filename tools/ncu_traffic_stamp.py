#!/usr/bin/env python3
"""Records the DRAM traffic of a k_prune_tc5 capture in profiles/ncu_prune_traffic.json, stamped with the hash of the kernel sources.
usage: tools/ncu_traffic_stamp.py profiles/ncu_prune_tc5_<tag>_traffic.json <columns per captured launch> <summary file>"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
src, cols, summary = sys.argv[1], int(sys.argv[2]), sys.argv[3]
t = json.load(open(src))
p = os.path.join(ROOT, "profiles", "ncu_prune_traffic.json")
d = json.load(open(p))
d.update(dram_bytes_per_launch_tc5=t["dram_bytes_per_launch"], tc5_source=summary, tc5_columns_per_captured_launch=cols, tc5_kernel_sha16=bench.kernel_sha16())
json.dump(d, open(p, "w"), indent=1)
print(json.dumps(d))
