#!/bin/bash
mkdir -p gpurun_out
items=""
for n in n16384 n65536; do for mb in 24 40 56 72 96 128 256; do items="$items $n:::FSEA_CHUNK_MB=$mb"; done; done
python tools/sweep.py $items > gpurun_out/sweep26.txt 2>&1
cat gpurun_out/sweep26.txt
