#!/bin/bash
# ncu launch list of the default bench workload (cold-cache, serialised per-launch times: shares only)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-3000} -c ${COUNT:-600} --csv \
  --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-target \
  > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"; tail -2 gpurun_out/ncu_launches.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_c2.csv')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter(); seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split('(')[0].replace('void folp::', '').replace('folp::', '')
    v = float(r[mv].replace(',', ''))
    if r[mu] == 'ns': v /= 1000
    tot[name] += v; cnt[name] += 1; seq.append((name, v))
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:56]:56s} n={cnt[k]:4d} total {v:9.1f} us avg {v/cnt[k]:8.2f} share {100*v/T:5.1f}%")
i = [j for j, (n, _) in enumerate(seq) if 'k_make_avg' in n]
if i:
    for n, v in seq[max(0, i[0]-2):i[0] + 16]: print(f"   {n[:50]:50s} {v:8.2f}")
PY
