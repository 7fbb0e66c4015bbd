#!/bin/bash
cp plz4_b200/libplz4cu.so /tmp/keep.so
for v in build/variants/*.so; do
  cp $v plz4_b200/libplz4cu.so; touch plz4_b200/libplz4cu.so
  echo -n "$(basename $v): "; PLZ4CU_DEC_DUO=1 timeout 200 python bench.py --gib 4 --steps 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('decompress', d['decompress_gbs'])"
  PLZ4CU_DEC_DUO=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none \
      -k regex:lz4_decompress -c 1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "dram__|inst_executed|duration|issue_active" | awk '{printf "    %s %s %s\n", $1, $2, $3}'
done
cp /tmp/keep.so plz4_b200/libplz4cu.so
