#!/bin/bash
# swap pre-built library variants in and time the bench kernels with each (build/variants/[a-s]*.so)
cp plz4_b200/libplz4cu.so /tmp/keep.so
for v in build/variants/[a-s]*.so; do
  cp $v plz4_b200/libplz4cu.so; touch plz4_b200/libplz4cu.so
  echo -n "$(basename $v): "; timeout 200 python bench.py --gib 4 --steps 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('compress', d['compress_gbs'], 'decompress', d['decompress_gbs'])"
done
cp /tmp/keep.so plz4_b200/libplz4cu.so
