#!/bin/bash
for pdl in 0 1; do for nb in 2 3; do
  BSQ_PDL=$pdl BSQ_NB=$nb python tools/sweep_bf.py 2>&1 | tail -1
done; done
BSQ_PDL=1 BSQ_NB=2 BSQ_TMA_CTAS=3 python tools/sweep_bf.py 2>&1 | tail -1
BSQ_PDL=1 BSQ_NB=3 BSQ_TMA_CTAS=3 python tools/sweep_bf.py 2>&1 | tail -1
