#!/bin/bash
# A/B on the beam-only harness: each argument is "name:ENV=VAL,ENV=VAL" (env applied to one prof_pair run)
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  ( IFS=,; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    POB_PROF_UNIQUE=1024 timeout 300 python tools/prof_pair.py ${AB_PAIRS:-2664} 2 2>&1 | tail -2 | sed "s/^/[$name] /" ) | tee -a gpurun_out/ab_${AB_TAG:-x}.txt
done
