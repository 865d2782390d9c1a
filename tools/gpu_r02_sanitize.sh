#!/bin/bash
# compute-sanitizer passes over one small run of every kernel family; logs kept under profiles/.
set -u
O=gpurun_out; mkdir -p $O
export IU_GRAPH=0
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > $O/r02_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|^ok' $O/r02_sanitizer_$tool.txt | tr '\n' ' ')"
done
