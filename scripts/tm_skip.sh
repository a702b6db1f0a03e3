#!/bin/bash
# skip experiments on the tensor-memory-operand convolution: which side of the pipeline bounds a tile
for d in 0 1 3 4 8 12 15; do
  echo "== DMB_TM_DBG=$d"
  DMB_TM_DBG=$d timeout 120 python -u scripts/tm_layers.py 8192 2>&1 | awk '{print $1, $2, $3, $4, $6, $7, $8}'
done
