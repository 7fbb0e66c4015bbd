#!/bin/bash
# Round-end GPU pass on one B200: the whole GPU suite, the default bench line, the reference arm, ncu captures of both
# bench kernels and the launch list.  Everything lands in gpurun_out/ (summaries are copied to profiles/ by hand).
tag=${1:-r02f}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/pytest_gpu_$tag.txt
python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
bash tools/ncu_round.sh $tag c d l > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.txt; cut -c1-700 gpurun_out/bench_n1_$tag.json; cut -c1-400 gpurun_out/bench_ref_$tag.json; ls -la gpurun_out/*$tag*
