#!/bin/bash
# per pre-built library variant: bench speed of both kernels, then DRAM bytes / instructions of one decompress launch
cp plz4_b200/libplz4cu.so /tmp/keep.so
mkdir -p gpurun_out
for v in build/variants/*.so; do
  cp $v plz4_b200/libplz4cu.so; touch plz4_b200/libplz4cu.so
  echo -n "$(basename $v): "; timeout 200 python bench.py --gib 4 --steps 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('compress', d['compress_gbs'], 'decompress', d['decompress_gbs'])"
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none \
      -k regex:lz4_decompress_kernel -c 1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "dram__|inst_executed|duration|hit_rate" | awk '{printf "    %s %s %s\n", $1, $2, $3}'
done
cp /tmp/keep.so plz4_b200/libplz4cu.so
