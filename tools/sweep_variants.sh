#!/bin/bash
# development: kernel times of the loop body for every variants/*.so build (CHEST_LIB override)
for lib in variants/*.so; do
  printf "%-26s " $lib
  CHEST_LIB=$PWD/$lib python tests/gpu_quick_timing.py ${1:-4096} 2 2>&1 | grep "^batch" | tail -1 | sed 's/.*realizations.s \([0-9.]*\).*one_tap.: \([0-9.]*\).*total.: \([0-9.]*\).*k_ic_main.: \([0-9.]*\).*k_ic_light.: \([0-9.]*\).*perfect_csi_chain.: \([0-9.]*\).*/real\/s \1 one_tap \2 total \3 main \4 light \5 chain \6/'
done
