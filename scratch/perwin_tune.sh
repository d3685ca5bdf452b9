#!/bin/bash
# tuning sweep of the per-window float kernel (SDRG_EXPERIMENTS build): K x tile limit x stages x groups
export PYTHONPATH=.
for cfg in "64,15" "96,15" "128,20" "32,15" "64,32"; do
  for k in 1 2 4; do for st in 3 2; do for g in 4 2; do
    r=$(SDRG_FOLD_PERWIN_K=$k SDRG_FOLD_PERWIN_STAGES=$st SDRG_FOLD_PERWIN_GROUPS=$g timeout 120 python scratch/float_sweep.py $cfg 2>&1 | grep kernel | sed 's/.*= \([0-9.]*\) GB.*/\1/')
    echo "$cfg K=$k stages=$st groups=$g : $r"
  done; done; done
done
