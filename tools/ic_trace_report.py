"""Development tool: stage timeline of one persistent k_ic launch from the CHEST_IC_TRACE dump.
usage: python tools/ic_trace_report.py trace.bin"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8)
t = t[t[:, 1] > 0]
ts = t[:, 1:6].astype(np.int64)
t0 = ts[:, 0].min()
ts = (ts - t0) / 1e3
print("CTAs %d on %d SMs" % (len(t), len(set(t[:, 0].astype(int)))))
print("pre  stage ends  (grid barrier passed): %.1f us" % ts[:, 1].max())
print("main stage: own work done  min %.1f  median %.1f  max %.1f us ; barrier passed %.1f us" % (
    ts[:, 2].min(), np.median(ts[:, 2]), ts[:, 2].max(), ts[:, 3].max()))
print("post stage ends: %.1f us" % ts[:, 4].max())
print("units per CTA in main: min %d max %d" % (t[:, 7].min(), t[:, 7].max()))
