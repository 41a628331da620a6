import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import chest_b200
from oracle.ds import DSConfig, ds_setup
from tests.helpers import context_from_oracle
B, I = 1024, 4
S = ds_setup(DSConfig())
ctx = context_from_oracle(S, max_batch=B)
ctx.generate_draws(B, 1, 10 ** 9)
host = ctx.download_draws(B)
st, keep = ctx.pack_draws(host)
pinned = {}
for k, a in keep.items():
    pinned[k] = torch.from_numpy(a.view(np.float64) if a.dtype == np.complex128 else a).pin_memory()
    print(k, pinned[k].is_pinned(), pinned[k].numel() * pinned[k].element_size() / 1e6, "MB")
st.doppler_u, st.phase_u, st.noise = pinned["du"].data_ptr(), pinned["pu"].data_ptr(), pinned["noise"].data_ptr()
for name, sid in chest_b200.context.SCHEME_ID.items():
    if "b" + name in pinned: st.bits[sid] = pinned["b" + name].data_ptr()
for key, wid in (("pil_idx_fbmc", 0), ("pil_idx_ofdm", 1)):
    if key in pinned: st.pilot_idx[wid] = pinned[key].data_ptr()
err = torch.zeros((B, 7, I + 1, 3, 2, 2), dtype=torch.int32).pin_memory()
# raw H2D bandwidth
d = torch.empty_like(pinned["noise"], device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): d.copy_(pinned["noise"], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("raw pinned H2D GB/s", 5 * pinned["noise"].numel() * 8 / dt / 1e9)
def direct(n):
    for _ in range(n):
        ctx._check(ctx.lib.chest_run_batch(ctx._h, B, I, C.byref(st), 0, 0, C.c_void_p(err.data_ptr())))
def pref(n):
    dev = ctx.prefetch_draws(B, st)
    for i in range(n):
        nxt = ctx.prefetch_draws(B, st) if i + 1 < n else None
        ctx._check(ctx.lib.chest_run_batch(ctx._h, B, I, C.byref(dev), 0, 0, C.c_void_p(err.data_ptr())))
        dev = nxt
for name, f in (("direct", direct), ("prefetch", pref), ("direct", direct), ("prefetch", pref)):
    f(2)
    t = time.perf_counter(); f(6); dt = time.perf_counter() - t
    print(name, "ms/step", 1e3 * dt / 6)
def resident(n):
    for i in range(n): ctx.run_batch_device(B, I, None, seed=1, first_rep=i * B, err_dev_ptr=None)
