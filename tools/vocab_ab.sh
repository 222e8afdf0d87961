#!/bin/bash
# GPU: A/B timing of the vocabulary-statistics kernel variants inside the headline decode
for v in online o2 plain; do
  echo "== CAPB200_VOCAB_STATS=$v"
  CAPB200_VOCAB_STATS=$v timeout 300 python tools/kernel_table.py 2>&1 | grep -E "total kernel time|vocab_stats"
done
