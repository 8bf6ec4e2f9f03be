O=gpurun_out/r02j; mkdir -p $O
timeout 400 python tools/sweep.py n32768 n65536 cfg5 n32768::build_alt/libfsea_stas.so n65536::build_alt/libfsea_stas.so n32768:::FSEA_WS=0 n65536:::FSEA_WS=0 > $O/sweep.txt 2>&1; cat $O/sweep.txt
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 1 -c 1 -o $O/prof python tools/sweep.py n65536@2048 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/prof.ncu-rep > $O/ncu_full.txt 2>&1; python tools/ncu_ops.py $O/prof.ncu-rep > $O/ncu_ops.txt 2>&1
ncu -i $O/prof.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_source.csv.gz
rm -f $O/prof.ncu-rep
cat $O/ncu_full.txt
