#!/usr/bin/env python
"""Stamps profiles/traffic_ncu.json from an `ncu --set full` report of the default bench workload:

  python tools/traffic_stamp.py gpurun_out/full_c2.ncu-rep c2 profiles/r02_ncu_full_c2_summary.csv

dram__bytes_read.sum + dram__bytes_write.sum per launch of k_spmv<EpiDual> and k_spmv<EpiTrans>
(average over the captured launches), summed -- bench.py's roofline.traffic -- together with the
hash of the kernel sources the report was taken from (bench.py drops the number when it is stale).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep, workload, summary = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def to_bytes(name, r):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    per = {}
    for r in rows[2:]:
        k = r[col["Kernel Name"]]
        per.setdefault(k, []).append(to_bytes("dram__bytes_read.sum", r) + to_bytes("dram__bytes_write.sum", r))
    dual = [v for k, vs in per.items() if "EpiDual" in k for v in vs]
    trans = [v for k, vs in per.items() if "EpiTrans" in k for v in vs]
    if not dual or not trans:
        raise SystemExit(f"no k_spmv launches in {rep}: {list(per)}")
    traffic = sum(dual) / len(dual) + sum(trans) / len(trans)
    path = os.path.join(ROOT, "profiles", "traffic_ncu.json")
    try:
        d = json.load(open(path))
    except Exception:
        d = {}
    d[workload] = {"traffic_bytes": traffic, "k_spmv_EpiDual_bytes": sum(dual) / len(dual),
                   "k_spmv_EpiTrans_bytes": sum(trans) / len(trans), "launches_averaged": [len(dual), len(trans)],
                   "kernel_source_sha16": bench.kernel_source_hash(), "from": summary}
    json.dump(d, open(path, "w"), indent=1)
    print(json.dumps(d[workload]))


if __name__ == "__main__":
    main()
