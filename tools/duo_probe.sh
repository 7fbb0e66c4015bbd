#!/bin/bash
# the one-warp-per-block decoder against the two-warps-per-block one (PLZ4CU_DEC_DUO=1): tests, speed, DRAM traffic
for duo in 0 1; do
  echo "== PLZ4CU_DEC_DUO=$duo"
  [ $duo = 1 ] && PLZ4CU_DEC_DUO=1 timeout 600 python -m pytest tests/test_gpu_decompress.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
  PLZ4CU_DEC_DUO=$duo timeout 200 python bench.py --gib 4 --steps 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('decompress', d['decompress_gbs'])"
  PLZ4CU_DEC_DUO=$duo timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none \
      -k regex:lz4_decompress -c 1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "dram__|inst_executed|duration|hit_rate|issue_active" | awk '{printf "    %s %s %s\n", $1, $2, $3}'
done
