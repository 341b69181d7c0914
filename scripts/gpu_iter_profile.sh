#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
# launch list of the whole process; the solve is the tail: aggregate the last launches by kernel name
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_iter_launches.csv python scripts/iter_profile.py > gpurun_out/${tag}_iter.log 2>&1
tail -2 gpurun_out/${tag}_iter.log
python - <<PY
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/${tag}_iter_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
data = rows[1:]
# keep the launches after the last IluFactorColour (= the Krylov solve)
last = max(i for i, r in enumerate(data) if "IluFactor" in r[ki])
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data[last + 1:]:
    m = re.search(r"kernel1d(?:Pf)?<(?:dab::)?([A-Za-z0-9_]+)", r[ki])
    name = m.group(1) if m else re.sub(r"\(.*", "", r[ki])
    t = float(r[vi].replace(",", "")); t = t / 1000.0 if r[ui] in ("ns", "nsecond") else (t if r[ui] in ("us", "usecond") else t * 1000.0)
    agg[name][0] += 1; agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print("total device us in the solve phase: %.0f" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-28s n=%5d  %9.0f us  %5.1f %%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
