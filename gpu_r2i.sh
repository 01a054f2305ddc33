#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python bench.py --mode train --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_train.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_train.log | cut -c1-300
python - <<'PY'
import csv, collections, re
rows=[]
with open('gpurun_out/train_launches.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
for row in r:
    try: rows.append((row['Kernel Name'], float(row['Metric Value'].replace(',','')), row.get('Metric Unit','')))
    except Exception: pass
print(len(rows),'launches')
# last step = last quarter roughly: find the last occurrence of the optimizer kernel pair
idx=[i for i,(n,_,_) in enumerate(rows) if 'sgd_apply' in n]
print('sgd_apply at', idx[-6:])
if len(idx)>=2:
    seg=rows[idx[-2]+1: idx[-1]+1]
    tot=collections.Counter(); cnt=collections.Counter()
    for n,v,u in seg:
        scale = 1e-3 if u in ('ns','nsecond') else (1.0 if u in ('us','usecond') else 1e3)
        k=re.sub(r'<.*','',n)[:60]
        tot[k]+=v*scale; cnt[k]+=1
    print('last step: %d kernels, %.2f ms total' % (len(seg), sum(tot.values())/1e3))
    for k,v in tot.most_common(28): print('%9.1f us %5d  %s' % (v, cnt[k], k))
PY
