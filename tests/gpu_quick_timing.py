"""Development script (not a test): stage timings of the loop body on a B200 with oracle-built
setup inputs.  python tests/gpu_quick_timing.py [batch] [reps] [fp64|split_bf16] [product|oracle]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.ds import DSConfig, ds_setup
from tests.helpers import context_from_oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
precision = sys.argv[3] if len(sys.argv) > 3 else "fp64"
setup = sys.argv[4] if len(sys.argv) > 4 else "product"      # product: the library's own setup (modem description included)
estimator = sys.argv[5] if len(sys.argv) > 5 else "auto"
t = time.time()
if setup == "product":
    from chest_b200.simulation import DoublySelectiveSimulation
    sim = DoublySelectiveSimulation(max_batch=B, seed=1)
    ctx = sim.ctx
else:
    S = ds_setup(DSConfig())
    ctx = context_from_oracle(S, max_batch=B)
print("context setup s", time.time() - t)
print("fp64 peak dmma TF/s", ctx.fp64_peak("dmma"), "dfma TF/s", ctx.fp64_peak("dfma"))
print("work model", ctx.work_model(4))
ctx.set_precision(precision)
ctx.set_estimator_mode(estimator)
ctx.set_profiling(True)
for i in range(reps):
    t = time.time()
    err = ctx.run_batch(B, 4, None, seed=1, first_rep=i * B)
    dt = time.time() - t
    print("batch", B, "wall s", round(dt, 4), "realizations/s", round(B / dt, 1), {k: round(v, 3) for k, v in ctx.stage_times().items()},
          {k: round(v, 3) for k, v in ctx.kernel_times().items()})
nb = ctx.bit_counts()
ber = err.astype(float).mean(axis=0)
for sid, n in enumerate(("aux", "cod", "ofdm")):
    print(n, "BER est it0..4 @40dB", np.round(ber[-1, :, sid, 0, 0] / nb[sid, 0], 4),
          "perfect", np.round(ber[-1, :, sid, 1, 0] / nb[sid, 0], 4))
print("launches", ctx.launch_count(), "precision", ctx.precision_info())
print("estimator", {n: ctx.estimator_info(n) for n in ("aux", "cod", "ofdm")})
