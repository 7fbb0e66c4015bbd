#!/bin/bash
# reader throughput (2 GiB frame of 64 KiB blocks) against the read-ahead depth; three processes of three passes each
for ra in ${*:-3 4 5 6}; do
  echo -n "read_ahead $ra:"
  for rep in 1 2 3; do PLZ4CU_READ_AHEAD=$ra timeout 200 python tools/read_prof.py 2048 4 2>&1 | grep "^read" | tail -2 | awk '{printf " %s", $2}'; done; echo
done
