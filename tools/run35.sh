#!/bin/bash
# frame-per-unit shapes A/B for 4096 / 8192 (second batch)
mkdir -p gpurun_out
python tools/sweep.py n4096::build_alt/libfsea_z.so n4096::build_alt/libfsea_p.so n4096::build_alt/libfsea_q.so \
   n8192::build_alt/libfsea_z.so n8192::build_alt/libfsea_p.so n8192::build_alt/libfsea_q.so n8192::build_alt/libfsea_r.so \
   cfg3::build_alt/libfsea_z.so cfg3::build_alt/libfsea_p.so cfg3::build_alt/libfsea_q.so cfg3::build_alt/libfsea_r.so > gpurun_out/sweep35b.txt 2>&1
cat gpurun_out/sweep35b.txt
