#!/bin/bash
# Runs on the GPU box (under gpurun): a set of GPU parity tests, one bench line and the parity report.
# usage: bash profiles/gpu_check.sh <tag> "<pytest paths>" [bench args]
TAG=${1:-chk}
TESTS=${2:-tests}
OUT=gpurun_out/r2
mkdir -p $OUT
if [ "$TESTS" != "none" ]; then
  timeout 1200 python -m pytest -m gpu $TESTS -x -q -s > $OUT/${TAG}_test.log 2>&1
  tail -4 $OUT/${TAG}_test.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${3} > $OUT/${TAG}.json 2> $OUT/${TAG}.err
tail -3 $OUT/${TAG}.err
python profiles/parity_report.py > $OUT/${TAG}_parity.log 2>&1
tail -2 $OUT/${TAG}_parity.log
