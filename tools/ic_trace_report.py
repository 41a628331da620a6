"""Development tool: per-CTA timeline of the last k_ic_main launch from the CHEST_IC_TRACE dump
(8 x uint64 per CTA: [0] SM id, [1] start, [3] own work done (globaltimer ns), [7] units processed).
usage: CHEST_IC_TRACE=trace.bin python tests/gpu_quick_timing.py 1024 1; python tools/ic_trace_report.py trace.bin"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8)
t = t[t[:, 1] > 0]
t0 = t[:, 1].astype(np.int64).min()
start = (t[:, 1].astype(np.int64) - t0) / 1e3
done = (t[:, 3].astype(np.int64) - t0) / 1e3
print("CTAs %d on %d SMs" % (len(t), len(set(t[:, 0].astype(int)))))
print("start  : max %.1f us" % start.max())
print("done   : min %.1f  median %.1f  max %.1f us" % (done.min(), np.median(done), done.max()))
print("units per CTA: min %d  max %d  total %d" % (t[:, 7].min(), t[:, 7].max(), t[:, 7].sum()))
