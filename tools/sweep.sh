#!/bin/bash
# usage: tools/sweep.sh "13 12 11" "0 1 2"   -> compress GB/s and ratio per (hash bits, lazy) on the bench workload
for hb in $1; do for lz in $2; do
PLZ4CU_HASH_BITS=$hb PLZ4CU_LAZY=$lz python bench.py --gib 2 --steps 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HB $hb LAZY $lz comp', d['compress_gbs'], 'decomp', d['decompress_gbs'], 'ratio', d['compressed_ratio'], 'vs liblz4 %.4f' % (d['compressed_ratio']/0.38178))"
done; done
