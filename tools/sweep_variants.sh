#!/bin/bash
# development: time k_ic stage durations for every build/lib_*.so variant
for lib in build/lib_*.so; do
  for mode in est_fbmc all; do
    if [ "$mode" = all ]; then unset CHEST_DEBUG_ONLY; else export CHEST_DEBUG_ONLY=$mode; fi
    CHEST_LIB=$PWD/$lib CHEST_IC_TRACE=gpurun_out/tr.bin python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,numpy as np
d=json.loads(sys.stdin.read())
t=np.fromfile('gpurun_out/tr.bin',dtype=np.uint64).reshape(-1,8); t=t[t[:,1]>0]; ts=(t[:,1:6].astype(np.int64)-t[:,1].min())/1e3
print('%-28s %-9s value %8.0f  ic/iter %6.2f ms | pre %5.2f main %6.2f post %5.2f ms' % ('$lib','$mode',d['value'],d['stage_ms_per_step']['ic_iterations']/4, ts[:,1].max()/1e3,(ts[:,3].max()-ts[:,1].max())/1e3,(ts[:,4].max()-ts[:,3].max())/1e3))"
  done
done
