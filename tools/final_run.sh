#!/bin/bash
# Round-end GPU pass: full GPU suite, bench line, team-kernel probes, secondary configs, launch list, one ncu capture of
# the team kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand).  usage: tools/final_run.sh [tag]
tag=${1:-r01h}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_$tag.txt
python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
( timeout 200 python tools/team_probe.py; PLZ4CU_TEAM=0 timeout 200 python tools/team_probe.py ) > gpurun_out/team_probe_$tag.txt 2>&1
timeout 600 python tools/bench_configs.py > gpurun_out/configs_$tag.json 2> gpurun_out/configs_$tag.err
TEAM_PROBE_QUICK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:team -c 1 \
    -o gpurun_out/prof_team_$tag -f python tools/team_probe.py > gpurun_out/prof_team_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/launches_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.txt; cat gpurun_out/bench_n1_$tag.json | cut -c1-400; cat gpurun_out/team_probe_$tag.txt
