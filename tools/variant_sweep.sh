#!/bin/bash
# A/B the analysis kernel variants (PQMF_VARIANT), 3 runs each to see run-to-run noise
for v in 0 1 2 4 8 15 0; do
  for r in 1 2; do
    echo -n "V=$v : "; PQMF_VARIANT=$v timeout 100 python tools/quick_bench.py 2>&1 | grep "exact=False" | sed 's/synthesis.*//'
  done
done
