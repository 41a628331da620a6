#!/bin/bash
# development: stage times of the loop body for every variants/*.so build (CHEST_LIB override)
for lib in variants/*.so; do
  printf "%-22s " $lib
  CHEST_LIB=$PWD/$lib python tests/gpu_quick_timing.py ${1:-1024} 3 2>&1 | grep "^batch" | tail -1 | sed 's/.*realizations.s \([0-9.]*\).*k2_transmission_matrix.: \([0-9.]*\).*one_tap.: \([0-9.]*\).*ic_iterations.: \([0-9.]*\).*total.: \([0-9.]*\).*/real\/s \1  k2 \2  one_tap \3  ic \4  total \5/'
done
