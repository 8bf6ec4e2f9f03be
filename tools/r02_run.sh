O=gpurun_out/r02q; mkdir -p $O
timeout 400 python tools/sweep.py n16384 n16384::build_alt/libfsea_gather.so cfg4 cfg4::build_alt/libfsea_gather.so n16384 n16384::build_alt/libfsea_gather.so > $O/sweep.txt 2>&1; cat $O/sweep.txt | cut -c1-45,140-250
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_step or large_batch or sweep or capture" > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
SWEEP_STEPS=1 SWEEP_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fsea_ -s 1 -c 1 -o $O/prof python tools/sweep.py n16384@8192 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/prof.ncu-rep > $O/ncu_full.txt 2>&1; python tools/ncu_ops.py $O/prof.ncu-rep > $O/ncu_ops.txt 2>&1
ncu -i $O/prof.ncu-rep --page source --csv 2>/dev/null | gzip > $O/ncu_source.csv.gz
rm -f $O/prof.ncu-rep
cat $O/ncu_full.txt
