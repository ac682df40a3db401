#!/usr/bin/env python
"""Reads probe_kernels.py JSON lines on stdin and prints one compact line per build."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "lib" not in d:
        print(line.strip()[:300])
        continue
    print((d["lib"] + " " + ",".join(f"{k}={v}" for k, v in d.get("env", {}).items())).ljust(40),
          "K1 %.1f K2 %.1f K3 %.1f iter %.1f plainA %.1f plainAt %.1f run %.0f pure %.0f" % (
              d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"], d["plain_A_us"],
              d["plain_At_us"], d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
