O=gpurun_out/r02s; mkdir -p $O
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 1 -c 1 -o $O/prof python tools/sweep.py n32768@4096 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/prof.ncu-rep > $O/ncu_full.txt 2>&1; python tools/ncu_ops.py $O/prof.ncu-rep > $O/ncu_ops.txt 2>&1
ncu -i $O/prof.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_source.csv.gz
rm -f $O/prof.ncu-rep
cat $O/ncu_full.txt
