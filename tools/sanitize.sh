#!/bin/bash
# compute-sanitizer over the default-path kernels (race / memory checking).  Run on the GPU box from the repository root:
#     bash tools/sanitize.sh [pcg] [step] [slab]          (slab needs gpurun --gpus 2)
# Logs: gpurun_out/sanitize_<tool>_<target>.log; the summaries kept under profiles/ are the tails of those logs.
# Eager launches (BLUB_NO_GRAPH=1): the sanitizer instruments kernels launched through graphs as well, but reports are easier to attribute this way.
set -u
OUT=gpurun_out
mkdir -p $OUT
TARGETS="${*:-pcg step}"
export BLUB_NO_GRAPH=1
for target in $TARGETS; do
    for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
        log=$OUT/sanitize_${tool}_${target}.log
        timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py $target > $log 2>&1
        echo "== $tool $target: rc $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
        grep -E "^(pcg|step|slab) " $log | tail -12
    done
done
