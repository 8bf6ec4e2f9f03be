#!/bin/bash
mkdir -p gpurun_out
for fr in 128 256 384 512 768; do
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k regex:fsea_ -s 12 -c 4 --csv --log-file gpurun_out/l2test_$fr.csv python tools/sweep.py n16384@$fr > gpurun_out/l2test_$fr.log 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/l2test_*.csv'), key=lambda s:int(s.split('_')[-1].split('.')[0])):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    d={}
    for r in rows:
        d.setdefault((r[0],r[4][:40]),{})[r[-3]]=r[-1]
    print(f)
    for k,v in d.items(): print('  ',k[1], {a:b for a,b in v.items()})
PY
