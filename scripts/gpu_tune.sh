#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune2.log
for sw in 0 1; do
for mb in 1024 4096 16384 65536; do
    echo "== XMP_SWEEP=$sw XMP_MIN_BATCH=$mb" >> gpurun_out/tune2.log
    XMP_SWEEP=$sw XMP_MIN_BATCH=$mb python scripts/time_search.py mh1 Metazoan mh3 h4 mh4 rbc14x759 EukaryotesrDNA 2>&1 | grep rep1 | cut -c1-16,95-260 >> gpurun_out/tune2.log
done
done
cat gpurun_out/tune2.log
