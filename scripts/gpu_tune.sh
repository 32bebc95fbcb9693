#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tune5.log
for gb in 4 12 48; do
    echo "== XMP_BUDGET_GB=$gb" >> gpurun_out/tune5.log
    XMP_BUDGET_GB=$gb python scripts/time_search.py mh1 Metazoan mh3 h4 mh4 EukaryotesrDNA 2>&1 | grep rep1 | cut -c1-16,40-260 >> gpurun_out/tune5.log
done
cat gpurun_out/tune5.log
