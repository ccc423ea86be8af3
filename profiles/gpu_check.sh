#!/bin/bash
# usage (under gpurun): bash profiles/gpu_check.sh <tag> [bench args]  -> gpurun_out/pytest_<tag>.log, bench_<tag>.json
tag=$1; shift
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 300 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cat gpurun_out/bench_$tag.json; tail -4 gpurun_out/bench_$tag.err
