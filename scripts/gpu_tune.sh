#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune3.log
for fb in 0 1; do
for mb in 4096 32768 131072; do
    echo "== XMP_SWEEP_FALLBACK=$fb XMP_MIN_BATCH=$mb" >> gpurun_out/tune3.log
    XMP_SWEEP_FALLBACK=$fb XMP_MIN_BATCH=$mb python scripts/time_search.py mh1 Metazoan mh3 h4 mh4 rbc14x759 EukaryotesrDNA mh2 2>&1 | grep rep1 | cut -c1-16,95-260 >> gpurun_out/tune3.log
done
done
cat gpurun_out/tune3.log
