#!/bin/bash
# swap pre-built library variants in and time the team decoder with each (build/variants/t_*.so against a_*.so);
# TEAM_COPY=n passes PLZ4CU_TEAM_COPY to the t_ variants
cp plz4_b200/libplz4cu.so /tmp/keep.so
for v in build/variants/a_*.so build/variants/t_*.so; do
  cp $v plz4_b200/libplz4cu.so; touch plz4_b200/libplz4cu.so
  echo "$(basename $v):"
  case $(basename $v) in t_*) export PLZ4CU_TEAM_COPY=${TEAM_COPY:-16};; *) unset PLZ4CU_TEAM_COPY;; esac
  TEAM_PROBE_QUICK=1 timeout 120 python tools/team_probe.py 2>&1 | tail -1
done
cp /tmp/keep.so plz4_b200/libplz4cu.so
