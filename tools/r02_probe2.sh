#!/bin/bash
# one source-level capture of a sweep.py workload: r02_probe2.sh <workload@frames> <outdir>
W=$1; O=gpurun_out/$2; mkdir -p $O
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 1 -c 1 -o $O/prof python tools/sweep.py $W > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/prof.ncu-rep > $O/ncu_full.txt 2>&1
python tools/ncu_ops.py $O/prof.ncu-rep > $O/ncu_ops.txt 2>&1
ncu -i $O/prof.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_source.csv.gz
ncu -i $O/prof.ncu-rep --page details 2>/dev/null > $O/ncu_details.txt
rm -f $O/prof.ncu-rep
ls -la $O
