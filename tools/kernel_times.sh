#!/bin/bash
# per-kernel total device time of one benchmark step under ncu (duration-only pass), for a list of libraries
#   tools/kernel_times.sh <kernel regex> lib1.so lib2.so ...
rx=$1; shift
for lib in "$@"; do
  ALR_LIBRARY=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$rx -s $((${SKIP:-39})) -c ${COUNT:-13} --csv --log-file /tmp/kt.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python - "$lib" <<'PY'
import csv, sys, collections
lines=[l for l in open('/tmp/kt.csv') if not l.startswith('==')]
rows=list(csv.reader(lines)); hdr=rows[0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(float); n=collections.Counter()
for r in rows[1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',','')); v={'ns':v/1e6,'us':v/1e3,'ms':v}.get(r[ui], v)
    agg[r[ki].split('(')[0]]+=v; n[r[ki].split('(')[0]]+=1
print(sys.argv[1], {k:(round(v,3), n[k]) for k,v in agg.items()})
PY
done
