#!/bin/bash
# usage: tools/sweep_jump.sh "32 40 48 64 96"   -> compress GB/s and ratio per jump threshold on the bench workload
for j in $1; do
PLZ4CU_JUMP=$j python bench.py --gib 2 --steps 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('JUMP $j comp', d['compress_gbs'], 'decomp', d['decompress_gbs'], 'ratio', d['compressed_ratio'], 'vs liblz4 %.4f' % (d['compressed_ratio']/0.38178))"
done
