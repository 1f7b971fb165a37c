#!/bin/bash
# A/B timing of library builds inside ONE gpurun call: tools/ab.sh "<args of tools/prof_run.py>" name1 name2 ...
# (name = gpurun_scratch/lib_<name>.so, or "prod" for the in-tree library); every library runs twice, interleaved.
args="$1"; shift
for round in 1 2; do
  for n in "$@"; do
    if [ "$n" = "prod" ]; then lib=""; else lib="$PWD/gpurun_scratch/lib_$n.so"; fi
    echo "== $n (round $round)"
    CONMECH_B200_LIB=$lib CM_PROF_REPS=${CM_PROF_REPS:-5} python tools/prof_run.py $args 2>/dev/null | grep -v "^  "
  done
done
