#!/bin/bash
# swap pre-built library variants in and time the compress kernel with each (build/variants/*.so)
cp plz4_b200/libplz4cu.so /tmp/keep.so
for v in build/variants/*.so; do
  cp $v plz4_b200/libplz4cu.so; touch plz4_b200/libplz4cu.so
  echo -n "$(basename $v): "; timeout 100 python tools/cta_probe.py 2 3 | tail -1
done
cp /tmp/keep.so plz4_b200/libplz4cu.so
