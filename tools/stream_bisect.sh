#!/bin/bash
for s in 1 2 4 8; do
  echo "== sites $s"; PLZ4CU_COPY_SITES=$s timeout 300 python tools/stream_probe.py 512 2>&1 | grep -E "stream cx=0|Assertion" | head -2
done
