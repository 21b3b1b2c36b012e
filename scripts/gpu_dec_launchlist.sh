mkdir -p gpurun_out
for mode in mma skip cuda; do
  if [ $mode == skip ]; then export TEO_DEC_DBG_SKIP=1; else unset TEO_DEC_DBG_SKIP; fi
  if [ $mode == cuda ]; then export TEO_DEC_ATTN=cuda; else unset TEO_DEC_ATTN; fi
  timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/dec_ll_$mode.csv python tools/dec_attn_bench.py > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/dec_ll_$mode.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[hi+2:]:
    if len(r)>mv: d[r[kn][:50]].append(float(r[mv].replace(',','')))
for k,v in d.items():
    if len(v)>5: print('$mode', k, 'n',len(v),'median us', sorted(v)[len(v)//2]/1000)
PY
done
