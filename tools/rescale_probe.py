#!/usr/bin/env python
"""Wall-clock of rescale_problem (src/preprocess.jl:631-687, CLI defaults: Ruiz 10 + Pock-Chambolle
alpha 1) on a bench workload: the device path (folp_rescale_problem, upload and download included),
the CPU oracle's C restatement (one thread, as the reference) and the NumPy host mirror."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from folp_b200 import preprocess  # noqa: E402
from folp_b200.lib import rescale_problem as device_rescale  # noqa: E402
from folp_b200.synthetic import random_sparse_lp  # noqa: E402
from oracle import oracle  # noqa: E402

n, m, k = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
lp = random_sparse_lp(n, m, k)
out = {}
for name, fn in (("device", lambda: device_rescale(10, False, 1.0, lp)),
                 ("device (2nd call)", lambda: device_rescale(10, False, 1.0, lp)),
                 ("oracle C, 1 thread", lambda: oracle.rescale_problem(10, False, 1.0, lp)),
                 ("numpy host mirror", lambda: preprocess.rescale_problem(10, False, 1.0, 0, lp))):
    t0 = time.perf_counter()
    r = fn()
    out[name] = (time.perf_counter() - t0, r)
    print(f"rescale_problem {name:22s} {out[name][0] * 1e3:9.1f} ms", flush=True)
a, b = out["device"][1], out["oracle C, 1 thread"][1]
print("bit-identical to the oracle:", np.array_equal(a.scaled_qp.constraint_matrix.data, b.scaled_qp.constraint_matrix.data)
      and np.array_equal(a.constraint_rescaling, b.constraint_rescaling)
      and np.array_equal(a.variable_rescaling, b.variable_rescaling))
