#!/bin/bash
for cfg in "2048 64" "1024 32" "512 16" "256 16" "256 8"; do
  set -- $cfg
  echo "== chunk blocks $1 min MiB $2"
  PLZ4CU_CHUNK_BLOCKS=$1 PLZ4CU_CHUNK_MIB=$2 timeout 300 python tools/stream_probe.py 1024 2>&1 | grep -E "stream cx=0|batch_host pinned" | head -4
done
