#!/usr/bin/env python
"""Times folp_create's host half (no CUDA call) on a bench workload: six calls in one process."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from folp_b200.lib import host_prepare_ms  # noqa: E402

lp, params, holder, fparams, scaled = bench.make_problem(sys.argv[1] if len(sys.argv) > 1 else "c2")
print("host prepare ms (FOLP_THP=%s, %d cores):" % (os.environ.get("FOLP_THP"), os.cpu_count()),
      [round(host_prepare_ms(holder)) for _ in range(6)])
