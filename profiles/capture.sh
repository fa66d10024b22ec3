#!/bin/bash
# Runs on the GPU box (under gpurun): one `ncu --set full` capture of the heavy kernels of the second
# whole-chain step, exported as CSV pages (the .ncu-rep itself can exceed gpurun's 64 MiB return limit).
# usage: bash profiles/capture.sh <tag> [kernel regex] [launches to skip] [launches to capture] ["kernels whose source page to export"]
TAG=${1:-cap}
REGEX=${2:-'d4c_body|refine_|channel_kernel|response_kernel|ct_frame|ps_sequential|harvest_tail|lt_frame|rng_fill'}
OUT=gpurun_out
mkdir -p $OUT
N=$(python - <<PY
import re
print(len(re.findall(r'\|', r'''$REGEX''')) + 1)
PY
)
SKIP=${3:-14}
COUNT=${4:-14}
ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -s $SKIP -c $COUNT -o /tmp/prof_$TAG \
    python profiles/run_step.py 3 > $OUT/cap_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>> $OUT/cap_$TAG.log
for k in ${5:-d4c_body refine_ channel_kernel response_kernel}; do
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:$k > $OUT/src_${TAG}_$k.csv 2>/dev/null
done
SZ=$(stat -c %s /tmp/prof_$TAG.ncu-rep)
if [ "$SZ" -lt 30000000 ]; then cp /tmp/prof_$TAG.ncu-rep $OUT/; fi
ls -la $OUT /tmp/prof_$TAG.ncu-rep
