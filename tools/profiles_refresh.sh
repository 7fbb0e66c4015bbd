#!/bin/bash
# the secondary measurements of a round on one B200: streams (2 GiB), configs beside the CPU path, large blocks, team decoder
tag=${1:-r02f}
mkdir -p gpurun_out
( timeout 500 python tools/stream_probe.py 2048; timeout 300 python tools/stream_multi_probe.py 2>&1 | tail -3 ) > gpurun_out/streams_$tag.txt 2>&1
timeout 900 python tools/bench_configs.py > gpurun_out/configs_$tag.json 2> gpurun_out/configs_$tag.err
( timeout 200 python tools/team_probe.py ) > gpurun_out/team_$tag.txt 2>&1
tail -25 gpurun_out/streams_$tag.txt; head -c 3000 gpurun_out/configs_$tag.json; tail -3 gpurun_out/configs_$tag.err; cat gpurun_out/team_$tag.txt
