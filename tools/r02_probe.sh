#!/bin/bash
# Round-2 first contact: full ncu captures (with source) of the kernels to be reworked.
O=gpurun_out/r02a
mkdir -p $O
for w in cfg3 n16384@8192 cfg4@8192 n65536@2048 n1024@262144; do
  tag=$(echo $w | tr '@' '_')
  WL=$w; if [ "$w" = "n1024@262144" ]; then WL=cfg2; fi
  SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 1 -c 1 -o $O/prof_$tag \
     python tools/sweep.py $WL > $O/ncu_$tag.log 2>&1
  python tools/ncu_summary.py $O/prof_$tag.ncu-rep > $O/ncu_full_$tag.txt 2>&1
  python tools/ncu_ops.py $O/prof_$tag.ncu-rep > $O/ncu_ops_$tag.txt 2>&1
  ncu -i $O/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_source_$tag.csv.gz
  ncu -i $O/prof_$tag.ncu-rep --page details 2>/dev/null > $O/ncu_details_$tag.txt
  rm -f $O/prof_$tag.ncu-rep
done
ls -la $O
