#!/bin/bash
# usage: tools/sweep_env.sh VAR "v1 v2 ..."   -> compress GB/s and ratio per value of an engine tuning variable
for j in $2; do
env $1=$j python bench.py --gib 2 --steps 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1=$j comp', d['compress_gbs'], 'decomp', d['decompress_gbs'], 'ratio', d['compressed_ratio'], 'vs liblz4 %.4f' % (d['compressed_ratio']/0.38178))"
done
