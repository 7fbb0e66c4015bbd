#!/bin/bash
# ncu captures of the bench workload (8 GiB of 64 KiB blocks): one --set full report per kernel, then the launch list.
# usage: tools/ncu_round.sh tag [c|d|l ...]   (default: all three)   -> gpurun_out/prof_<tag>_{c,d}.ncu-rep, launches_<tag>.csv
tag=$1; shift; what=${*:-c d l}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    c) timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_compress_cta_kernel -c 1 -o gpurun_out/prof_${tag}_c -f \
         python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_${tag}_c.log 2>&1 ;;
    d) timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_kernel -c 1 -o gpurun_out/prof_${tag}_d -f \
         python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_${tag}_d.log 2>&1 ;;
    l) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
         python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_${tag}.log 2>&1 ;;
  esac
done
ls -la gpurun_out/prof_${tag}_* gpurun_out/launches_${tag}.csv 2>/dev/null
