"""Development script: where does the split-BF16 mode deviate from the FP64 mode?  python tools/tc_diag.py [B] [n_iter]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.ds import DSConfig, ds_setup
from tests.helpers import context_from_oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = ds_setup(DSConfig())
ctx = context_from_oracle(S, max_batch=B)
names = ("aux", "cod", "ofdm")
def grab():
    return {(n, r, s, w): ctx.get_state(w, n, r, s) for n in names for r in range(B) for s in (0, ctx.n_snr - 1) for w in ("xD_est", "hP", "hdiag")}
ref_err = ctx.run_batch(B, n_iter, None, seed=5, first_rep=1000); ref = grab()
ctx.set_precision("split_bf16")
got_err = ctx.run_batch(B, n_iter, None, seed=5, first_rep=1000); got = grab()
for n in names:
    for s in (0, ctx.n_snr - 1):
        for w in ("hP", "hdiag", "xD_est"):
            dev = np.array([np.max(np.abs(ref[(n, r, s, w)] - got[(n, r, s, w)])) / np.max(np.abs(ref[(n, r, s, w)])) for r in range(B)])
            worst = int(np.argmax(dev))
            d = np.abs(ref[(n, worst, s, w)] - got[(n, worst, s, w)])
            print("%-5s snr %d %-7s median %.2e max %.2e at rep %d index %d (n above 1e-4: %d of %d reps; entries above 1e-4 in worst rep: %d of %d)"
                  % (n, s, w, np.median(dev), dev.max(), worst, int(np.argmax(d)), int((dev > 1e-4).sum()), B,
                     int((d > 1e-4 * np.max(np.abs(ref[(n, worst, s, w)]))).sum()), d.size))
print("counter cells that differ:", int((ref_err != got_err).sum()), "of", ref_err.size)
