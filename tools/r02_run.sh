O=gpurun_out/r02o; mkdir -p $O
timeout 400 python tools/sweep.py n16384 n16384::build_alt/libfsea_late.so cfg4 cfg4::build_alt/libfsea_late.so n16384 n16384::build_alt/libfsea_late.so > $O/sweep.txt 2>&1; cat $O/sweep.txt | cut -c1-45,140-240
