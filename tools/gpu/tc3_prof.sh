#!/bin/bash
# where the split-precision step spends its time: ncu launch list of one step + the library's in-stream kernel timers
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc3.csv python bench.py --math tc3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu_tc3.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_tc3.csv')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
i_n, i_v, i_id = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
body = [r for r in rows if r is not hdr and r[i_id].isdigit()]
# last step only: launches after the final k_canonicalize
last = max(k for k, r in enumerate(body) if 'k_canonicalize' in r[i_n])
agg = collections.OrderedDict()
for r in body[last:]:
    k = r[i_n].split('(')[0][-60:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[i_v].replace(',', ''))
tot = sum(v[1] for v in agg.values())
for k, (n, t) in agg.items(): print(f"{k:62s} {n:3d} {t/1e3:9.1f} us {t/tot:6.1%}")
print('sum', tot / 1e3, 'us')
PY
