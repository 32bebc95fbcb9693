#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune.log
for pol in 0 1 2; do
for mb in 4096 32768 131072; do
  for beam in 64 512; do
    echo "== XMP_POLICY=$pol XMP_MIN_BATCH=$mb XMP_BEAM=$beam" >> gpurun_out/tune.log
    XMP_POLICY=$pol XMP_MIN_BATCH=$mb XMP_BEAM=$beam python scripts/time_search.py mh1 Metazoan mh3 h4 mh4 rbc14x759 2>&1 | grep rep1 | cut -c1-16,95-260 >> gpurun_out/tune.log
  done
done
done
cat gpurun_out/tune.log
