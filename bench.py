#!/usr/bin/env python
"""bench.py -- headline benchmark of the Monte-Carlo hot path (DS.m:350-565).

  python bench.py --gpus N --steps K --warmup W            our arm (B200, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm (oracle restatement of the reference)

metric  : channel realizations/sec (est+IC+BER)            (BASELINE.json)
workload: DoublySelectiveChannelEstimation.m with its default parameters: N = 540 samples,
          OFDM (K = 336) + FBMC auxiliary symbols + FBMC data spreading (K = 720), 16 pilots,
          7 SNR points, 4 interference-cancellation iterations, estimated- and perfect-CSI chains.
step    : one batch of `--batch` realizations per GPU through the whole loop body.
value   : realizations/s with every input resident in HBM (draws generated on the device inside
          the timed region, results left on the device).
e2e     : the same through the host-facing call chest_run_batch with HOST buffers: explicit draws
          copied host->device from pinned memory and error counts copied back, every step.
Prints exactly one JSON line on rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "channel realizations/sec (est+IC+BER)"
UNIT = "realizations/s"
WORKLOAD_PAPER = ("DoublySelectiveChannelEstimation.m with the paper block DS.m:42-46 enabled: N=7350, OFDM K=672 + FBMC-Aux + "
                  "FBMC-Cod K=1440, P=32, 6 taps, 16 SNR points (10:2:40 dB), 4 IC iterations, estimated + perfect CSI")
WORKLOAD = ("DoublySelectiveChannelEstimation.m default parameters (DS.m:16-37): N=540, OFDM K=336 + FBMC-Aux + "
            "FBMC-Cod K=720, P=16, 7 SNR points, 4 IC iterations, estimated + perfect CSI, 24 BER outputs")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="realizations per step per GPU")
    ap.add_argument("--schemes", default="aux,cod,ofdm")
    ap.add_argument("--no-ofdm-only", action="store_true", help="skip the extra OFDM-chain-only measurement")
    ap.add_argument("--no-split-leg", action="store_true", help="skip the extra leg in the split-BF16 tensor-core mode")
    ap.add_argument("--no-dense-leg", action="store_true", help="skip the extra leg with D materialised (K2 rooflines)")
    ap.add_argument("--cpu-sample", type=int, default=48, help="realizations timed for cpu_baseline (about 13 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-sample", action="store_true", help="skip the oracle check of sampled realizations")
    ap.add_argument("--workload", default="default", choices=["default", "paper", "sweep", "sv", "scaled"],
                    help="default: DS.m default parameters (BASELINE.json configs[1], the headline); paper: DS.m:42-46 "
                         "(configs[2]); sweep: velocity sweep, 1e5 realizations over all GPUs including setup (configs[3]); "
                         "sv: SimpleVersion_DoublyFlat.m chain with the FFT modem on the device (configs[0])")
    ap.add_argument("--sweep-realizations", type=int, default=100000)
    ap.add_argument("--scaled-subcarriers", type=int, default=300, help="--workload scaled: 300 / 600 / 1200 (5 / 10 / 20 MHz LTE grids)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_realizations_per_s(n_sample, schemes, faithful=False):
    """Oracle restatement of DS.m:350-565 timed on the host cores (cpu_baseline / reference arm)."""
    import numpy as np
    from oracle.ds import DSConfig, ds_setup, ds_realization, new_draws
    S = ds_setup(DSConfig(schemes=tuple(schemes)))
    rng = np.random.default_rng(0)
    draws = [new_draws(S, rng) for _ in range(n_sample + 1)]
    ds_realization(S, draws[-1], faithful=faithful)            # warm-up
    t = time.perf_counter()
    for d in draws[:n_sample]:
        ds_realization(S, d, faithful=faithful)
    return n_sample / (time.perf_counter() - t)


def dense_mode_line(torch, Simulation, schemes, device, B, K, W, I, peak_dmma, hbm_peak, hbm_src):
    """Same workload, perfect-CSI pass in dense mode (chest_set_perfect_csi_mode(DENSE)): K2 on the timed path."""
    sim = Simulation(schemes=tuple(schemes), max_batch=B, device=device, seed=1234)
    ctx = sim.ctx
    ctx.set_perfect_csi_mode("dense")
    ctx.set_profiling(True)
    err_dev = torch.zeros(B * len(sim.Pn) * (I + 1) * 12, dtype=torch.int32, device="cuda")
    step = [0]

    def run():
        ctx.run_batch_device(B, I, None, seed=sim.seed, first_rep=step[0] * B, err_dev_ptr=err_dev.data_ptr())
        step[0] += 1
    for _ in range(W):
        run()
    torch.cuda.synchronize()
    ctx.event_record(0)
    kern, hg_ms, hg_bytes = {}, 0.0, 0.0
    for _ in range(K):
        run()
        for k, v in ctx.kernel_times().items():
            kern[k] = kern.get(k, 0.0) + v
        a, b = ctx.banded_apply_stats()
        hg_ms, hg_bytes = hg_ms + a, hg_bytes + b
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    wm = ctx.work_model(I)
    k2_ms = kern["k_gemm_d"] / K
    ic_ms = kern["k_ic_main"] / (K * I)
    ic_flops = B * (wm["est_main_flops"] + wm["perf_flops"]) / I
    line = {"value": B * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "realizations_per_step": B,
            "kernel_ms_per_step": {k: v / K for k, v in kern.items()},
            "roofline_k_ic_main": {"bound": "tensor", "achieved": ic_flops / (ic_ms * 1e-3) / 1e12, "peak": peak_dmma,
                                   "unit": "TFLOP/s", "frac": ic_flops / (ic_ms * 1e-3) / 1e12 / peak_dmma,
                                   "avg_launch_ms": ic_ms},
            "roofline_k2": {"kernel": "k_gemm_d (D = Q^H H G, persistent, support-aware)", "bound": "tensor",
                            "achieved": B * wm["k2_flops"] / (k2_ms * 1e-3) / 1e12, "peak": peak_dmma, "unit": "TFLOP/s",
                            "frac": B * wm["k2_flops"] / (k2_ms * 1e-3) / 1e12 / peak_dmma,
                            "algorithmic_flops_per_launch": B * wm["k2_flops"], "avg_launch_ms": k2_ms},
            "roofline_k1": {"kernel": "k_apply_hg (banded, never-materialised H applied to G)", "bound": "hbm",
                            "achieved": hg_bytes / (hg_ms * 1e-3) / 1e9 if hg_ms > 0 else None, "peak": hbm_peak,
                            "unit": "GB/s", "frac": hg_bytes / (hg_ms * 1e-3) / 1e9 / hbm_peak if hg_ms > 0 else None,
                            "peak_source": hbm_src, "algorithmic_bytes_per_step": hg_bytes / K, "ms_per_step": hg_ms / K}}
    sim.close()
    return line


def split_bf16_line(torch, np, ctx, sim, B, K, W, I, err_fp64_last, last_first, wm):
    """The stated reduced-precision mode on the same context and workload (chest_set_precision(SPLIT_BF16)): the
    estimated-CSI cancellation runs on tcgen05 tensor cores (k_ic_est_tc: split-BF16 operands, FP32 accumulators in TMEM),
    everything else stays FP64.  Reports the step time, the tensor roofline of the kernel against the measured BF16 peak,
    and how many hard decisions differ from the FP64 mode on the last timed batch."""
    n_snr = len(sim.Pn)
    ctx.set_precision("split_bf16")
    err_dev = torch.zeros(B * n_snr * (I + 1) * 12, dtype=torch.int32, device="cuda")
    step = [10 ** 6]

    def run(first=None):
        ctx.run_batch_device(B, I, None, seed=sim.seed, first_rep=step[0] * B if first is None else first,
                             err_dev_ptr=err_dev.data_ptr())
        step[0] += 1
    for _ in range(W):
        run()
    torch.cuda.synchronize()
    ctx.event_record(0)
    kern = {}
    for _ in range(K):
        run()
        for k, v in ctx.kernel_times().items():
            kern[k] = kern.get(k, 0.0) + v
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    run(first=last_first)                                       # the batch the FP64 mode processed last
    torch.cuda.synchronize()
    got = err_dev.view(B, n_snr, I + 1, 3, 2, 2).cpu().numpy().astype(np.int64)
    ref = err_fp64_last.astype(np.int64)
    nb = ctx.bit_counts()
    n_dec = B * n_snr * I * float(sum(nb[sid, 0] for sid in range(3)))
    mode, mma_flops, img_bytes = ctx.precision_info()
    ctx.set_precision("fp64")
    tc_ms = kern["k_ic_main"] / (K * I)
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(pk["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step; burst %.0f)" % pk["bf16_tflops"]
    except Exception:
        peak, peak_src = 1400.0, "B200_PROFILING.md fallback, sustained 1.4 PFLOP/s (MEASURED_PEAKS.json absent)"
    alg = B * wm["est_main_flops"] / I
    return {"value": B * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "realizations_per_step": B, "dtype": "bf16x2 split operands, f32 accumulate (estimated-CSI cancellation only; the rest f64)",
            "kernel_ms_per_step": {k: v / K for k, v in kern.items()},
            "roofline": {"kernel": "k_ic_est_tc (tcgen05.mma kind::f16 BF16, 128 x 256 x 16, three products per k-step, TMEM accumulators, bulk-copied W images)",
                         "bound": "tensor", "achieved": mma_flops / (tc_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": mma_flops / (tc_ms * 1e-3) / 1e12 / peak, "peak_source": peak_src, "avg_launch_ms": tc_ms,
                         "executed_dense_bf16_flops_per_launch": mma_flops, "algorithmic_flops_per_launch": alg,
                         "algorithmic_tflops": alg / (tc_ms * 1e-3) / 1e12,
                         "note": "achieved counts the dense BF16 flops the MMAs execute: 3 split products x 4 real products per complex one x "
                                 "the block-dense tiles (128 rows x the union of their active columns; about 27 % of those entries are "
                                 "non-zero at the default geometry).  algorithmic_tflops is the 8-flop-per-complex-MAC count of the FP64 "
                                 "roofline over the same launch time, for comparison with k_ic_main"},
            "operand_image_bytes": int(img_bytes),
            "vs_fp64_mode": {"bit_decisions_that_differ": int(np.abs(got - ref)[:, :, 1:, :, 0, 0].sum()), "of": int(n_dec),
                             "perfect_csi_and_one_tap_counters_identical": bool(np.array_equal(got[:, :, :, :, 1], ref[:, :, :, :, 1]) and np.array_equal(got[:, :, 0], ref[:, :, 0])),
                             "tolerance": "estimated channel within 1e-4 of the FP64 mode (tests/test_gpu_tc.py measures 3e-6 on the pilot estimates, 1e-5 on diag(D_est))"}}


def factored_estimator_line(torch, np, ctx, sim, B, K, W, I, err_fp64_last, last_first, hbm_peak, run_e2e=None, h2d=0, d2h=0):
    """The stated-tolerance mode chest_set_estimator_mode(FACTORED) on the same context and workload: the estimated-CSI
    cancellation (D_est - diag h_est) v applied as Modulation -> estimated banded channel H_est = sum_q g_q M_q -> Demodulation
    (k_est_channel + k_est_factored) instead of through the thresholded W tiles.  FP64 arithmetic throughout; D_est differs
    from the reference's by the entries its 1e-8 thresholds removed.  Reports the step time, the kernel time of the factored
    pass, its HBM figure, and how many hard decisions differ from the default (tile-form) mode on the last timed batch."""
    n_snr = len(sim.Pn)
    ctx.set_estimator_mode("factored")
    err_dev = torch.zeros(B * n_snr * (I + 1) * 12, dtype=torch.int32, device="cuda")
    step = [2 * 10 ** 6]

    def run(first=None):
        ctx.run_batch_device(B, I, None, seed=sim.seed, first_rep=step[0] * B if first is None else first,
                             err_dev_ptr=err_dev.data_ptr())
        step[0] += 1
    for _ in range(W):
        run()
    torch.cuda.synchronize()
    ctx.event_record(0)
    kern, ef_ms = {}, 0.0
    for _ in range(K):
        run()
        for k, v in ctx.kernel_times().items():
            kern[k] = kern.get(k, 0.0) + v
        ef_ms += ctx.estimator_info(next(iter(ctx.schemes)))["ms"]
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    info = {n: ctx.estimator_info(n) for n in ctx.schemes}
    e2e = None
    if run_e2e is not None:                                     # the same end-to-end leg as the main line, in this mode
        run_e2e(2)
        torch.cuda.synchronize()
        t_e = time.perf_counter()
        run_e2e(K)
        torch.cuda.synchronize()
        e2e = {"value": B * K / (time.perf_counter() - t_e), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "chest_prefetch_draws + chest_run_batch with pinned host buffers, estimator mode FACTORED"}
    run(first=last_first)
    torch.cuda.synchronize()
    got = err_dev.view(B, n_snr, I + 1, 3, 2, 2).cpu().numpy().astype(np.int64)
    ref = err_fp64_last.astype(np.int64)
    nb = ctx.bit_counts()
    n_dec = B * n_snr * I * float(sum(nb[sid, 0] for sid in range(3)))
    ctx.set_estimator_mode("auto")
    n_cols = B * n_snr * sum(1 for n in ctx.schemes if info[n]["factored"])
    N_s, T_taps = ctx.N, ctx.T
    Kmean = sum(s_["K"] for s_ in ctx.schemes.values()) / max(1, len(ctx.schemes))
    # per column and launch: H_est written and read (2 x 16 T N), v + h_est + y read, y_ic written (4 x 16 K)
    bytes_launch = n_cols * (2.0 * 16 * T_taps * N_s + 4.0 * 16 * Kmean)
    pass_ms = ef_ms / (K * I)
    return {"value": B * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "realizations_per_step": B, "dtype": "f64", "e2e": e2e,
            "estimator": {n: {"factored": info[n]["factored"], "largest_removed_R_Dij_hP": info[n]["removed_r"],
                              "largest_removed_W": info[n]["removed_w"]} for n in info},
            "kernel_ms_per_step": dict({k: v / K for k, v in kern.items()}, est_factored_pass=ef_ms / K),
            "roofline": {"kernel": "k_est_channel + k_est_factored (H_est = sum_q g_q M_q per column; IFFT, filter / overlap-add, banded H_est, "
                                   "fold, FFT, cancellation epilogue in shared memory)", "bound": "hbm",
                         "achieved": bytes_launch / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_launch / (pass_ms * 1e-3) / 1e9 / hbm_peak if pass_ms > 0 else None, "avg_launch_ms": pass_ms,
                         "algorithmic_bytes_per_launch": bytes_launch, "columns_per_launch": n_cols,
                         "note": "the chain itself lives in shared memory (L1 data pipe 76 % busy, DRAM 31 %: "
                                 "profiles/r02_final_kernels_summary.txt); the HBM figure is reported because no other pipe has a stated peak"},
            "vs_default_mode": {"bit_decisions_that_differ": int(np.abs(got - ref)[:, :, 1:, :, 0, 0].sum()), "of": int(n_dec),
                                "perfect_csi_and_one_tap_counters_identical": bool(np.array_equal(got[:, :, :, :, 1], ref[:, :, :, :, 1]) and np.array_equal(got[:, :, 0], ref[:, :, 0])),
                                "tolerance": "D_est within 1e-4 of max|D_est| of the reference formulation (stated; tests/test_gpu_scale.py "
                                             "test_factored_estimator_stated_mode_fbmc bounds pilot, channel and symbol estimates and checks the "
                                             "counts against the oracle restatement of the factored form)"}}


def perfect_csi_roofline(ctx, sim, B, I, wm, chain_ms, peak_dmma, hbm_peak):
    """The perfect-CSI pass y - Q^H H G v + h v of one iteration.  FBMC and CP-OFDM columns go through the FFT-form modem
    (k_perfect_fbmc_det: G and Q^H applied as IFFT / filter / overlap-add -- or cyclic prefix -- and filter / fold / FFT, all in
    shared memory, with equalisation, detection and counters behind it): the pass is bound by the L1 data pipe (shared-memory and
    strided global wavefronts), not by HBM or the FP64 pipe, so its roofline is reported in bytes with the flops of the GEMM
    formulation it replaces beside it.  (Geometries without a usable modem description run k_gemm_ring x2 + k_apply_h_cols_planes.)"""
    n_snr = len(sim.Pn)
    cols = {name: B * n_snr for name in sim.sch}
    k_of = {name: ctx.schemes[name]["K"] for name in sim.sch}
    alg_bytes = sum(cols[n] * k_of[n] * 3 * 16.0 for n in sim.sch)           # v in, y in, y_ic out (16 B each per symbol)
    gemm_flops = B * wm["factored_perf_flops"] / I
    return {"kernel": "k_perfect_fbmc_det (FBMC and CP-OFDM columns: Modulation -> banded H -> Demodulation -> cancellation -> equalise / detect / count, "
                      "one column per CTA in shared memory; 24-point specialised chain at the default geometry)",
            "bound": "hbm", "achieved": alg_bytes / (chain_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / (chain_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_iteration": alg_bytes, "avg_iteration_ms": chain_ms,
            "gemm_formulation_flops_per_iteration": gemm_flops, "gemm_equivalent_tflops": gemm_flops / (chain_ms * 1e-3) / 1e12,
            "gemm_equivalent_frac_of_dmma_peak": gemm_flops / (chain_ms * 1e-3) / 1e12 / peak_dmma,
            "note": "algorithmic bytes: v and y read, y_ic written once per column (48 B per symbol); the kernel is bound by the L1 data pipe "
                    "(ncu l1tex 78 %, profiles/r02_final_kernels_summary.txt), not by HBM (11 %) or the FP64 pipe (16 %).  The GEMM formulation "
                    "of round 1 needed 13 x the flops; gemm_equivalent_* divides THOSE flops by the time of the pass (above 1 = faster than a "
                    "perfect DMMA GEMM could be)"}


def ofdm_only_line(torch, Simulation, device, K, W, cpu_sample):
    """realizations/s of the DS.m loop body with only the OFDM scheme enabled (K = 336), B = 4096 per step."""
    B, I = 4096, 4
    sim = Simulation(schemes=("ofdm",), max_batch=B, device=device, seed=1234)
    ctx = sim.ctx
    err_dev = torch.zeros(B * len(sim.Pn) * (I + 1) * 12, dtype=torch.int32, device="cuda")
    step = [0]

    def run():
        ctx.run_batch_device(B, I, None, seed=sim.seed, first_rep=step[0] * B, err_dev_ptr=err_dev.data_ptr())
        step[0] += 1
    for _ in range(W):
        run()
    torch.cuda.synchronize()
    ctx.event_record(0)
    for _ in range(K):
        run()
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    line = {"value": B * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "realizations_per_step": B,
            "workload": "DS.m default parameters with the OFDM chain only: N=540, K=336, P=16, 7 SNR points, 4 IC iterations"}
    sim.close()
    if cpu_sample:
        line["cpu_baseline"] = {"value": cpu_realizations_per_s(max(8, cpu_sample), ["ofdm"]), "unit": UNIT,
                                "cores": blas_threads(), "kind": "port"}
    return line


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_cores():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core it
    can (the launcher's setting would halve its value and inflate the ratio).  Returns the BLAS thread count in use."""
    import numpy as np  # noqa: F401  (loads BLAS so that threadpoolctl sees it)
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return blas_threads()


def parity_sample(schemes, seed, first, err_batch, n_iter, reps):
    """The checker, not the thing measured: realizations `reps` of the batch that the timed region processed last
    (counter-based draws keyed by (seed, first + r)) are re-run by the CPU oracle with ITS OWN setup; the GPU's error
    counts -- produced by the product's setup and kernels -- must equal the oracle's."""
    import numpy as np
    from oracle.ds import DSConfig, ds_setup, ds_realization
    from oracle import rng as orng
    from tests.helpers import err_from_oracle
    S = ds_setup(DSConfig(schemes=tuple(schemes), NrIterations=n_iter))
    bad = []
    for r in reps:
        ref = err_from_oracle(ds_realization(S, orng.draws_for(S, seed, first + r)), n_iter)
        if not np.array_equal(err_batch[r], ref):
            bad.append(int(first + r))
    return "ok" if not bad else "MISMATCH at realizations %s" % bad


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_cores()
    schemes = args.schemes.split(",")
    n = max(1, args.cpu_sample // 4)                      # a bounded sample per step: ~3 s of CPU work
    vals = []
    for _ in range(args.warmup):
        cpu_realizations_per_s(1, schemes)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_realizations_per_s(n, schemes))
    value = float(sum(vals) / len(vals))
    sample = ("oracle port (NumPy/OpenBLAS restatement of DS.m:350-565, support-aware D-hat as a GEMV; NOT MATLAB), "
              "%d realizations per step, %d steps" % (n, args.steps))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "schemes": schemes, "realizations_per_step": n},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
           "host_cores": os.cpu_count()}
    print(json.dumps(out), flush=True)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import chest_b200
    from chest_b200.simulation import DoublySelectiveSimulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    schemes = args.schemes.split(",")
    B, K, W, I = args.batch, args.steps, args.warmup, 4
    parity_failed = False
    paper = args.workload == "paper"
    workload = WORKLOAD
    if paper:
        if args.batch == 4096:
            B = 512
        workload = WORKLOAD_PAPER
    t0 = time.perf_counter()
    if paper:
        sim = DoublySelectiveSimulation.paper(schemes=tuple(schemes), max_batch=B, device=local, seed=1234)
    else:
        sim = DoublySelectiveSimulation(schemes=tuple(schemes), max_batch=B, device=local, seed=1234)
    setup_s = time.perf_counter() - t0
    ctx = sim.ctx
    n_snr = len(sim.Pn)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_dmma = ctx.fp64_peak("dmma", 20000)
    peak_dfma = ctx.fp64_peak("dfma", 20000)
    err_dev = torch.zeros(B * n_snr * (I + 1) * 12, dtype=torch.int32, device="cuda")
    step_id = [0]

    def step_resident():
        # shard: realization indices (step * world + rank) * B .. + B  (independent of the GPU count per index)
        first = (step_id[0] * world + rank) * B
        ctx.run_batch_device(B, I, None, seed=sim.seed, first_rep=first, err_dev_ptr=err_dev.data_ptr())
        step_id[0] += 1

    ctx.set_profiling(True)
    for _ in range(W):
        step_resident()
    clocks = ClockSampler(local)
    stage_sum = {}
    barrier()
    clocks.start()
    launches0 = ctx.launch_count()
    ctx.event_record(0)
    t_wall = time.perf_counter()
    hg_ms, hg_bytes = 0.0, 0.0
    kern_sum = {}
    for _ in range(K):
        step_resident()
        for k, v in ctx.stage_times().items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
        a, b = ctx.banded_apply_stats()
        hg_ms, hg_bytes = hg_ms + a, hg_bytes + b
        for k, v in ctx.kernel_times().items():
            kern_sum[k] = kern_sum.get(k, 0.0) + v
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    launches = ctx.launch_count() - launches0
    clk = clocks.stop()
    last_first = ((step_id[0] - 1) * world + rank) * B           # first realization index of the last timed batch
    err_last = err_dev.view(B, n_snr, I + 1, 3, 2, 2).cpu().numpy().astype(np.uint32)

    # ---- end-to-end leg: host draws (pinned) -> device, counts -> host, every step
    ctx.generate_draws(B, sim.seed, 10 ** 9)
    host = ctx.download_draws(B)
    st, keep = ctx.pack_draws(host)
    pinned = {}
    for k, a in keep.items():                                   # re-home the packed arrays in pinned memory
        t = torch.from_numpy(a.view(np.float64) if a.dtype == np.complex128 else a).pin_memory()
        pinned[k] = t
    st.doppler_u, st.phase_u, st.noise = pinned["du"].data_ptr(), pinned["pu"].data_ptr(), pinned["noise"].data_ptr()
    for name, sid in chest_b200.context.SCHEME_ID.items():
        if "b" + name in pinned:
            st.bits[sid] = pinned["b" + name].data_ptr()
    for key, wid in (("pil_idx_fbmc", 0), ("pil_idx_ofdm", 1)):
        if key in pinned:
            st.pilot_idx[wid] = pinned[key].data_ptr()
    err_host = torch.zeros((B, n_snr, I + 1, 3, 2, 2), dtype=torch.int32).pin_memory()
    h2d = ctx.draws_bytes(B)
    d2h = err_host.numel() * 4
    import ctypes as C

    # every step uploads its draws from pinned host memory and reads its counters back; the upload of step i+1 is
    # started (chest_prefetch_draws, copy stream) before step i is run, so it travels under step i's kernels
    def run_e2e(n_steps):
        dev = ctx.prefetch_draws(B, st)
        for i in range(n_steps):
            nxt = ctx.prefetch_draws(B, st) if i + 1 < n_steps else None
            rc = ctx.lib.chest_run_batch(ctx._h, B, I, C.byref(dev), 0, 0, C.c_void_p(err_host.data_ptr()))
            ctx._check(rc)
            dev = nxt

    run_e2e(max(2, W // 2))
    barrier()
    ctx.event_record(2)
    t_e = time.perf_counter()
    run_e2e(K)
    ctx.event_record(3)
    e2e_dev_ms = ctx.event_elapsed_ms(2, 3)
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t_e)
    launches_per_step = launches / K

    # ---- final reduce of the error counters over ranks (the only collective of the path)
    tot = err_dev.to(torch.int64).view(B, n_snr, I + 1, 12).sum(dim=0)
    t_max = torch.tensor([dev_ms, e2e_wall_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    dev_ms, e2e_wall_ms, wall_ms = [float(x) for x in t_max.cpu()]
    if rank == 0:
        wm = ctx.work_model(I)
        ic_ms = kern_sum["k_ic_main"] / (K * I)                            # average duration of one k_ic_main launch
        ic_flops = B * wm["est_main_flops"] / I                            # algorithmic flops of one launch (estimated-CSI units)
        chain_ms = kern_sum["perfect_csi_chain"] / (K * I)                 # G v, H, Q^H of one iteration, both waveforms
        nb = ctx.bit_counts()
        ber40 = {name: float(tot[-1, -1, sid * 4].item()) / float(nb[sid, 0] * B * world)
                 for name, sid in chest_b200.context.SCHEME_ID.items() if name in sim.sch}
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (MEASURED_PEAKS.json absent)"
        traffic, traffic_note = None, None
        try:                                                       # ncu --set full capture, scaled to this batch size
            tj = json.load(open(os.path.join(ROOT, "profiles", "kic_traffic.json")))
            traffic = tj["dram_bytes_per_launch"] * B / tj["batch"]
            traffic_note = "dram__bytes_read+write per k_ic_main launch from %s, scaled x%.2f to batch %d" % (
                tj["source"], B / tj["batch"], B)
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": world * B * K / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "schemes": schemes, "realizations_per_step_per_gpu": B,
                       "parallelism": "realizations sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "working set larger than L2 (per IC launch %.2f GB of MMSE tiles + %.1f GB of per-unit v / y_ic "
                             "columns vs 126 MB L2)" % (wm["w_bytes_per_ic_launch"] / 1e9, ctx.n_units() * 2 * 720 * 16 * 16 / 1e9),
                       "perfect_csi_mode": "factored (D never formed); dense_d_mode holds the leg with D materialised",
                       "timing": "CUDA events on the library's stream around all K steps; max over ranks",
                       "seed": sim.seed},
            "e2e": {"value": world * B * K / (e2e_wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "device_ms_per_step": e2e_dev_ms / K,
                    "wall_ms_per_step": e2e_wall_ms / K,
                    "api": "chest_prefetch_draws + chest_run_batch (C ABI) with pinned host buffers: explicit draws in (every "
                           "step, uploaded under the previous step's kernels), error counts out"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"kernel": "k_ic_main (one persistent launch per IC iteration: y - W_off(hP) v of every (scheme, SNR, realization) on FP64 DMMA)",
                         "bound": "tensor", "achieved": ic_flops / (ic_ms * 1e-3) / 1e12, "peak": peak_dmma,
                         "unit": "TFLOP/s", "frac": ic_flops / (ic_ms * 1e-3) / 1e12 / peak_dmma, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "peak_source": "measured in this run: register-resident FP64 DMMA (m8n8k4) loop on every SM "
                                        "(chest_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; DFMA probe %.1f TFLOP/s"
                                        % peak_dfma,
                         "algorithmic_flops_per_launch": ic_flops, "avg_launch_ms": ic_ms,
                         "frac_executed": 0.75 * ic_flops / (ic_ms * 1e-3) / 1e12 / peak_dmma,
                         "flops_note": "algorithmic flops: 8 per complex multiply-add (SURVEY.md 8d); the kernel executes "
                                       "the three-multiplication form (6 per complex multiply-add): frac can exceed 1, "
                                       "frac_executed = 0.75 frac is the share of the DMMA peak the executed products take "
                                       "(ncu pipe-busy fraction: profiles/r01_kic_main_summary.txt)"},
            "roofline_perfect_csi": perfect_csi_roofline(ctx, sim, B, I, wm, chain_ms, peak_dmma, hbm_peak),
            "stage_ms_per_step": {k: v / K for k, v in stage_sum.items()},
            "kernel_ms_per_step": {k: v / K for k, v in kern_sum.items()},
            "wall_ms_per_step": wall_ms / K, "launches_per_step": launches_per_step, "setup_s": setup_s,
            "sanity_ber_40dB_last_iteration": ber40,
        }
        # small kernels of the step (sincos/s of the synthesis, GB/s of the banded apply)
        T_taps, paths, N_s = ctx.T, ctx.paths, sim.N
        n_vec = sum(1 for n_ in sim.sch)                                   # transmit vectors per realization
        synth_ms, apply_ms = kern_sum["k_synth_h"] / K, kern_sum["k_apply_h"] / K
        demod_ms = stage_sum["k3_demod"] / K
        out["roofline_small_kernels"] = {
            "k_synth_h": {"bound": "fp64 alu (sincos)", "sincos_per_s": B * T_taps * paths * N_s / (synth_ms * 1e-3) if synth_ms > 0 else None,
                          "ms_per_step": synth_ms, "note": "T * paths * N complex exponentials per realization (FF.m:235); writes 16 T N bytes"},
            "k_apply_h": {"bound": "hbm", "achieved": B * n_vec * 16.0 * N_s * (T_taps + 2) / (apply_ms * 1e-3) / 1e9 if apply_ms > 0 else None,
                          "peak": hbm_peak, "unit": "GB/s",
                          "frac": B * n_vec * 16.0 * N_s * (T_taps + 2) / (apply_ms * 1e-3) / 1e9 / hbm_peak if apply_ms > 0 else None,
                          "ms_per_step": apply_ms, "note": "16 N (T + 2) algorithmic bytes per vector (read h, read s, write r), SURVEY.md 8d"},
            "k_gemm_demod": {"bound": "tensor", "achieved": B * wm["txdemod_flops"] / (demod_ms * 1e-3) / 1e12 if demod_ms > 0 else None,
                             "peak": peak_dmma, "unit": "TFLOP/s", "ms_per_step": demod_ms,
                             "note": "y = Q^H (r0 + noise) for every (scheme, SNR, realization); flops include the TX GEMM s = G x"},
            "k_tx_symbols": {"ms_per_step": kern_sum["k_tx_symbols"] / K}, "modulate_gemm": {"ms_per_step": kern_sum["modulate_gemm"] / K}}
        light_ms = kern_sum["k_ic_light"] / (K * (I + 1))
        light_bytes = ctx.n_units() * 2.0 * 16 * 16 * (sum(s_["K"] for s_ in ctx.schemes.values()) / max(1, len(ctx.schemes)))
        out["roofline_k_ic_light"] = {"kernel": "k_ic_light (LS pilots, W_diag x hP, equalise, decide, count, precode v)", "bound": "hbm",
                                      "achieved": light_bytes / (light_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": light_bytes / (light_ms * 1e-3) / 1e9 / hbm_peak, "avg_launch_ms": light_ms,
                                      "algorithmic_bytes_per_launch": light_bytes,
                                      "note": "y_ic read + v written once per unit (16 columns x K rows x 16 B each way); the kernel is "
                                              "instruction bound, not bandwidth bound: profiles/r02_kic_post_vs_light.txt"}
        if world == 1 and not args.no_split_leg:
            try:
                out["split_bf16_mode"] = split_bf16_line(torch, np, ctx, sim, B, K, W, I, err_last, last_first, wm)
            except Exception as e:                              # noqa: BLE001
                out["split_bf16_mode"] = {"error": repr(e)[:300]}
                try:
                    ctx.set_precision("fp64")
                except Exception:                               # noqa: BLE001
                    pass
        if world == 1 and not args.no_split_leg:
            try:
                out["factored_estimator_mode"] = factored_estimator_line(torch, np, ctx, sim, B, K, W, I, err_last, last_first, hbm_peak, run_e2e, h2d, d2h)
            except Exception as e:                              # noqa: BLE001
                out["factored_estimator_mode"] = {"error": repr(e)[:300]}
                try:
                    ctx.set_estimator_mode("auto")
                except Exception:                               # noqa: BLE001
                    pass
        if world == 1 and not args.no_dense_leg and not paper:
            # the same workload with D = Q^H H G materialised per realization (K2) and applied densely: the rooflines
            # of k_gemm_d (FP64 tensor) and k_apply_hg (HBM) come from this leg; never fatal for the main line
            try:
                out["dense_d_mode"] = dense_mode_line(torch, DoublySelectiveSimulation, schemes, local, min(B, 1024), K, W, I,
                                                      peak_dmma, hbm_peak, hbm_src)
            except Exception as e:                              # noqa: BLE001
                out["dense_d_mode"] = {"error": repr(e)[:200]}
        if world == 1 and set(schemes) != {"ofdm"} and not args.no_ofdm_only and not paper:
            # the same loop body with the OFDM chain alone (the narrow reading of "default params: OFDM"), for
            # comparison: resident value only, same timing rules; never fatal for the main line
            try:
                out["ofdm_chain_only"] = ofdm_only_line(torch, DoublySelectiveSimulation, local, K, W,
                                                        None if args.no_cpu_baseline else args.cpu_sample)
            except Exception as e:                              # noqa: BLE001
                out["ofdm_chain_only"] = {"error": repr(e)[:200]}
        if paper:
            out["parity_sample"] = "not run (the oracle's setup at the paper geometry takes minutes; parity at this geometry: tests/test_gpu_parity.py::test_paper_geometry_parity, tests/test_gpu_scale.py::test_reference_bundle_on_gpu[paper], ::test_figure5_statistical_pin)"
        elif not args.no_parity_sample:
            use_all_host_cores()
            reps = sorted({0, 17, B // 2 + 1, B - 1})
            out["parity_sample"] = parity_sample(schemes, sim.seed, last_first, err_last, I, reps)
            out["parity_sample_note"] = ("GPU error counts of realizations %s of the last timed batch (seed %d, first index %d) "
                                         "equal the CPU oracle's (own setup, same counter-based draws)" % (reps, sim.seed, last_first))
        if paper:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                                   "sample": "not timed at the paper geometry (oracle setup alone takes minutes); see the default workload"}
        elif world == 1 and not args.no_cpu_baseline:
            n = args.cpu_sample
            v = cpu_realizations_per_s(n, schemes)
            vf = cpu_realizations_per_s(1, schemes, faithful=True)
            out["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                "sample": "oracle port of DS.m:350-565 (NumPy/OpenBLAS, support-aware GEMV D-hat; not MATLAB), %d "
                          "realizations of the same workload" % n,
                "reference_faithful_value": vf,
                "reference_faithful_note": "same port with the reference's own formulation (dense Q'HG, full()+bsxfun "
                                           "D-hat, DS.m:388-389,417-425), 1 realization"}
        print(json.dumps(out), flush=True)
        if str(out.get("parity_sample", "ok")).startswith("MISMATCH"):
            sys.stderr.write("bench.py: parity sample failed: %s\n" % out["parity_sample"])
            parity_failed = True
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sim.close()
    if parity_failed:
        raise SystemExit(3)


def run_sweep(args):
    """BASELINE.json configs[3]: high-mobility sweep.  `--sweep-realizations` (1e5) realizations in total, split evenly over the
    velocities 50..500 km/h and over the GPUs, VehicularA, 4 IC iterations, default geometry, all three schemes -- timed
    INCLUDING the per-velocity setup (correlation + MMSE matrices rebuilt on the device, DS.m:208-313) and the final counter
    reduce; only the velocity-independent one-time setup (modem matrices, precoders) is outside.  Strong scaling."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from chest_b200.simulation import DoublySelectiveSimulation
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    velocities = list(range(50, 501, 50))
    I, B = 4, min(args.batch, 2048)
    per_v = args.sweep_realizations // len(velocities)
    t0 = time.perf_counter()
    sim = DoublySelectiveSimulation(max_batch=B, device=local, seed=77, Velocity_kmh=velocities[0])
    one_time_s = time.perf_counter() - t0
    sim.run(NrRepetitions=min(B, 256), NrIterations=I)                  # warm-up: allocations, first-launch costs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    setup_s, loop_s = 0.0, 0.0
    setup_list, setup_parts = [], []
    t_all = time.perf_counter()
    totals = []
    n_snr = len(sim.Pn)
    for v in velocities:
        t = time.perf_counter()
        sim.set_velocity(v)
        setup_s += time.perf_counter() - t
        setup_list.append(round(time.perf_counter() - t, 4)); setup_parts.append({k: round(x, 4) for k, x in sim.setup_times.items()})
        t = time.perf_counter()
        lo = per_v * rank // world; hi = per_v * (rank + 1) // world     # this rank's contiguous block of the velocity's realizations
        tot = sim.run_totals(hi - lo, NrIterations=I, first_rep=lo)
        totals.append(tot)
        loop_s += time.perf_counter() - t
    tt = torch.from_numpy(np.stack(totals).astype(np.int64)).cuda()
    t_red = time.perf_counter()
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)                       # the one collective: counters of all velocities
    torch.cuda.synchronize()
    reduce_ms = 1e3 * (time.perf_counter() - t_red)
    wall = time.perf_counter() - t_all
    times = torch.tensor([wall, setup_s, loop_s, reduce_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, setup_s, loop_s, reduce_ms = [float(x) for x in times.cpu()]
    if rank == 0:
        nb = sim.ctx.bit_counts()
        tt = tt.cpu().numpy().reshape(len(velocities), n_snr, I + 1, 3, 2, 2)
        n_total = per_v * len(velocities)
        ber = {str(v): float(tt[k, -1, I, 2, 0, 0]) / (nb[2, 0] * per_v) for k, v in enumerate(velocities)}
        out = {"metric": METRIC, "value": n_total / wall, "unit": UNIT, "n_gpus": world, "steps": len(velocities), "warmup": 1,
               "ms_per_step": 1e3 * wall / len(velocities), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": "high-mobility sweep (BASELINE.json configs[3]): velocities %s km/h, VehicularA, DS.m default geometry, "
                                      "3 schemes, 7 SNR points, 4 IC iterations, %d realizations in total over %d GPU(s), per-velocity "
                                      "setup (DS.m:208-313 on the device) inside the timed region" % (velocities, n_total, world),
                          "realizations_per_velocity": per_v, "batch": B, "timing": "host wall clock around the whole sweep (setup + loop + reduce), max over ranks"},
               "setup_s_per_velocity": setup_s / len(velocities), "loop_s": loop_s, "reduce_ms": reduce_ms, "one_time_setup_s": one_time_s,
               "loop_only_value": n_total / loop_s, "sanity_ber_ofdm_40dB_iteration4_by_velocity": ber,
               "setup_breakdown_last_velocity_s": {k: round(v, 4) for k, v in sim.setup_times.items()},
               "setup_s_by_velocity": setup_list, "setup_breakdown_by_velocity_s": setup_parts}
        if world == 1:
            # the same sweep in the stated-tolerance mode chest_set_estimator_mode(FACTORED): D_est = Q^H H_est G, FP64, without the
            # reference's 1e-8 thresholds; the per-velocity setup rebuilds the factors as well
            try:
                sim.ctx.set_estimator_mode("factored")
                sim.set_velocity(velocities[0]); sim.run_totals(min(B, 256), NrIterations=I, first_rep=0)       # warm-up of the new unit tables
                torch.cuda.synchronize()
                f_setup, f_loop, f_tot = 0.0, 0.0, []
                t_f = time.perf_counter()
                for v in velocities:
                    t = time.perf_counter(); sim.set_velocity(v); f_setup += time.perf_counter() - t
                    t = time.perf_counter(); f_tot.append(sim.run_totals(per_v, NrIterations=I, first_rep=0)); f_loop += time.perf_counter() - t
                torch.cuda.synchronize()
                f_wall = time.perf_counter() - t_f
                ft = np.stack(f_tot).astype(np.int64).reshape(len(velocities), n_snr, I + 1, 3, 2, 2)
                out["factored_estimator_mode"] = {
                    "value": n_total / f_wall, "unit": UNIT, "setup_s_per_velocity": f_setup / len(velocities), "loop_s": f_loop,
                    "loop_only_value": n_total / f_loop,
                    "factored": {n: sim.ctx.estimator_info(n)["factored"] for n in sim.sch},
                    "bit_decisions_that_differ_from_default_mode": int(np.abs(ft - tt)[:, :, 1:, :, 0, 0].sum()),
                    "of": int(n_total * n_snr * I * float(sum(nb[sid, 0] for sid in range(3))))}
                sim.ctx.set_estimator_mode("auto")
            except Exception as e:                              # noqa: BLE001
                out["factored_estimator_mode"] = {"error": repr(e)[:300]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sim.close()


def run_sv(args):
    """BASELINE.json configs[0]: SimpleVersion_DoublyFlat.m (L = 12, 30 FBMC / 15 OFDM symbols, 16-QAM / 4-PAM, SNR 0:5:30 dB,
    8 Diamond pilots, doubly-flat Rayleigh channel).  A "realization" of this workload is one body of the script's double loop,
    SV.m:91-170 (one repetition at one SNR point, all five BER outputs), run as batched device launches (chest_sv_run_batch)."""
    import numpy as np
    import torch
    from chest_b200.simulation import SimpleVersionSimulation
    from oracle.sv import sv_setup, sv_new_draws, sv_body, sv_pn
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(0)
    B, K, W = (8192 if args.batch == 4096 else args.batch), args.steps, args.warmup
    B -= B % 7                                                   # whole repetitions: 7 SNR points per repetition (SV.m:15)
    t0 = time.perf_counter()
    sim = SimpleVersionSimulation(max_batch=B, seed=1234)
    setup_s = time.perf_counter() - t0
    ctx = sim.ctx
    snrs = np.resize(np.asarray(sim.M_SNR_OFDM_dB, dtype=float), B)            # body = rep * nS + snr, as the script loops
    pn = sim.noise_power(snrs)
    step = [0]

    def run():
        err = ctx.sv_run_batch(pn, None, seed=sim.seed, first_body=step[0] * B)
        step[0] += 1
        return err
    for _ in range(W):
        run()
    clocks = ClockSampler(0)
    torch.cuda.synchronize()
    clocks.start()
    l0 = ctx.launch_count()
    ctx.event_record(0)
    t_w = time.perf_counter()
    for _ in range(K):
        err = run()
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms(0, 1)
    wall_ms = 1e3 * (time.perf_counter() - t_w)
    launches = ctx.launch_count() - l0
    clk = clocks.stop()
    # end to end: explicit host draws (bits, pilot indices, channel coefficient, noise of both waveforms) in, counts out
    S = sv_setup(sim.ChannelEstimation_FBMC.PilotMatrix, sim.AuxiliaryMethod.PilotMatrix, sim.ChannelEstimation_OFDM.PilotMatrix,
                 sim.interp_f, sim.interp_o)
    rng = np.random.default_rng(3)
    Be = min(B, 2048)
    base = [sv_new_draws(S, rng) for _ in range(64)]
    draws = [base[i % 64] for i in range(Be)]
    h2d = sum(np.asarray(v).nbytes for v in base[0].values()) * Be
    import ctypes as C
    from chest_b200 import _lib as clib
    from chest_b200.context import SCHEME_ID
    st, pinned = clib.ChestSvDraws(), {}

    def pin(key, arr):                                              # packed once, in pinned host memory
        a = np.ascontiguousarray(arr)
        pinned[key] = torch.from_numpy(a.view(np.float64) if a.dtype == np.complex128 else a).pin_memory()
        return pinned[key].data_ptr()
    for name, sid in SCHEME_ID.items():
        st.bits[sid] = pin("b" + name, np.stack([d["bits_" + name] for d in draws]).astype(np.uint8))
    for key, wid in (("pil_idx_fbmc", 0), ("pil_idx_ofdm", 1)):
        st.pilot_idx[wid] = pin(key, np.stack([d[key] for d in draws]).astype(np.int32))
    st.h = pin("h", np.array([d["h"] for d in draws], dtype=np.complex128))
    for key, wid in (("noise_fbmc", 0), ("noise_ofdm", 1)):
        st.noise[wid] = pin(key, np.stack([d[key] for d in draws]).astype(np.complex128))
    st.on_device = 0
    pn_e = np.ascontiguousarray(pn[:Be]); err_e = np.zeros((Be, 5), dtype=np.uint32)

    def run_e2e():
        ctx._check(ctx.lib.chest_sv_run_batch(ctx._h, Be, pn_e.ctypes.data, C.byref(st), 0, 0, err_e.ctypes.data))
    run_e2e()
    t_e = time.perf_counter()
    for _ in range(max(1, K // 2)):
        run_e2e()
    e2e_ms = 1e3 * (time.perf_counter() - t_e) / max(1, K // 2)
    # parity on the explicit draws: identical error counts to the oracle's restatement of SV.m:95-169
    bad = [b for b in range(0, 64, 7) if not np.array_equal(err_e[b], sv_body(S, draws[b], pn[b]))]
    # CPU baseline: the oracle port of the loop body on the host cores
    cores = use_all_host_cores()
    t_c = time.perf_counter(); n_c = 0
    while time.perf_counter() - t_c < 10.0:
        sv_body(S, base[n_c % 64], pn[n_c % len(pn)]); n_c += 1
    cpu = n_c / (time.perf_counter() - t_c)
    ber = (err.reshape(-1, len(sim.M_SNR_OFDM_dB), 5).astype(float).mean(axis=0) / sim.n_bits[None, :])
    N = sim.FBMC.Nr["SamplesTotal"]
    bytes_per_body = 16.0 * N * 2 * 2                               # transmit + receive signal of both waveforms (the FFT modem's traffic)
    out = {"metric": METRIC, "value": B * K / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "SimpleVersion_DoublyFlat.m default parameters (BASELINE.json configs[0]): L=12, 30 FBMC / 15 OFDM symbols, "
                                  "N=%d, 16-QAM / 4-PAM, SNR 0:5:30 dB, 8 Diamond pilots; one realization = one (repetition, SNR) body SV.m:91-170" % N,
                      "bodies_per_step": B, "timing": "CUDA events on the library's stream around all K steps", "seed": sim.seed,
                      "l2": "working set per step %.1f GB (signals of %d bodies) vs 126 MB L2" % (B * bytes_per_body / 1e9, B)},
           "e2e": {"value": Be / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(err_e.nbytes),
                   "bodies_per_step": Be, "api": "chest_sv_run_batch (C ABI) with explicit host draws in pinned memory: bits, pilot indices, h, noise of both waveforms in; error counts out"},
           "gpu_launches": int(launches), "clocks": clk, "wall_ms_per_step": wall_ms / K, "setup_s": setup_s,
           "roofline": {"kernel": "FFT modem + channel + detection chain (k_modem_ifft, k_fbmc_overlap_add, k_sv_channel, k_modem_fft, k_sv_detect)",
                        "bound": "hbm", "achieved": B * K * bytes_per_body / (dev_ms * 1e-3) / 1e9, "peak": _hbm_peak()[0], "unit": "GB/s",
                        "frac": B * K * bytes_per_body / (dev_ms * 1e-3) / 1e9 / _hbm_peak()[0], "traffic": None,
                        "note": "algorithmic bytes: the time-domain signals written by the modulators and read by the demodulators "
                                "(16 N bytes each way per waveform and body); the step is a chain of small launch-bound kernels, not one HBM-bound kernel",
                        "peak_source": _hbm_peak()[1]},
           "parity_sample": "ok" if not bad else "MISMATCH at bodies %s" % bad,
           "sanity_ber_by_snr": {n: [round(float(x), 5) for x in ber[:, k]] for k, n in enumerate(("FBMC_Aux", "FBMC_Cod", "FBMC_perfect", "OFDM", "OFDM_perfect"))},
           "cpu_baseline": {"value": cpu, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "oracle port of SV.m:95-169 (NumPy FFT modem), %d bodies in 10 s" % n_c}}
    print(json.dumps(out), flush=True)
    sim.close()
    if bad:
        raise SystemExit(3)


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (MEASURED_PEAKS.json absent)"


def run_scaled(args):
    """BASELINE.json configs[4]: scaled bandwidth.  L = 300 subcarriers (5 MHz LTE grid), fs = 512 x 15 kHz = 7.68 MHz, 30 FBMC
    symbols: N = 9472 samples, K = 9000 symbols -- the dense model of the reference (G, Q: 1.36 GB each; D: 1.3 GB per realization;
    H: N x N) does not fit a workstation, the banded H and the support-aware, chunked D GEMM do.  A step = B realizations of rows
    A-D of the hot path: NewRealization (sum of 200 sinusoids per tap), the banded H applied to G, D = Q^H (H G) with its diagonal."""
    import numpy as np
    import torch
    import chest_b200
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(0)
    L = args.scaled_subcarriers
    B, K, W = (8 if args.batch == 4096 else args.batch), args.steps, args.warmup
    nfft = 512 if L <= 300 else (1024 if L <= 600 else 2048)
    fs = 15e3 * nfft
    t0 = time.perf_counter()
    fb = chest_b200.Modulation.FBMC(L, 30, 15e3, fs, 15e3 * ((nfft - L) // 2), False, "Hermite-OQAM", 4, 0, True)
    N, Ksym = fb.Nr["SamplesTotal"], L * 30
    G = fb.GetTXMatrix()
    Q = fb.GetRXMatrix().conj().T
    ch = chest_b200.Channel.FastFading(fs, "VehicularA", N, 500 / 3.6 * 2.5e9 / 2.998e8, "Jakes", 200, 1, 1, False, create_device=False)
    ctx = chest_b200.DeviceContext(0)
    pdp = ch.Implementation["PowerDelayProfileNormalized"]
    ctx.set_channel(N, pdp, ch.PHY["MaximumDopplerShift"], ch.PHY["dt"], 200, "Jakes")
    ctx.set_waveform("F", G, Q)
    ctx.finalize(B)
    setup_s = time.perf_counter() - t0
    peak_dmma = ctx.fp64_peak("dmma", 20000)
    hbm_peak, hbm_src = _hbm_peak()
    T = ctx.T
    step = [0]

    def run():
        ctx.new_realization_seeded(B, 1234, step[0] * B)
        step[0] += 1
        return ctx.transmission_matrix_batch("F", B)
    for _ in range(W):
        run()
    clocks = ClockSampler(0)
    torch.cuda.synchronize()
    clocks.start()
    l0 = ctx.launch_count()
    ctx.event_record(0)
    hg_ms = gd_ms = 0.0
    for _ in range(K):
        a, b, flops = run()
        hg_ms += a; gd_ms += b
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms(0, 1)
    launches = ctx.launch_count() - l0
    clk = clocks.stop()
    # parity: sampled entries of D of the first and last realization against the host product with the device's own h
    rng = np.random.default_rng(1)
    taps = np.flatnonzero(pdp)
    worst = 0.0
    for b in (0, B - 1):
        h = ctx.impulse_response(b)                                  # N x Lt
        rows = rng.integers(0, Ksym, 48); cols = np.clip(rows + rng.integers(-2 * L, 2 * L + 1, 48), 0, Ksym - 1)
        rows[:8] = cols[:8]                                          # some diagonal entries
        got = ctx.transmission_matrix_entries("F", b, rows, cols)
        ref = np.zeros(48, dtype=complex)
        for e, (i, j) in enumerate(zip(rows, cols)):
            hg = np.zeros(N, dtype=complex)
            for m in taps:
                hg[m:] += h[m:, m] * G[:N - m, j]
            ref[e] = np.vdot(Q[:, i], hg)
        worst = max(worst, float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
    # end to end: explicit Doppler / phase draws from the host in, diag(D) of every realization out
    du = rng.random((B, T * 200)); pu = rng.random((B, T * 200))
    hd = torch.zeros((B, Ksym, 2), dtype=torch.float64).pin_memory()
    ctx.new_realization(du, pu); ctx.transmission_matrix_batch("F", B, hd.data_ptr())
    t_e = time.perf_counter()
    for _ in range(max(1, K // 2)):
        ctx.new_realization(du, pu); ctx.transmission_matrix_batch("F", B, hd.data_ptr())
    e2e_ms = 1e3 * (time.perf_counter() - t_e) / max(1, K // 2)
    # CPU baseline: the dense formulation of DS.m:388-389 on the host cores, on a slice of the columns of D
    cores = use_all_host_cores()
    import scipy.sparse as sp
    h0 = ctx.impulse_response(0)
    n_idx = np.arange(N)
    Hs = sp.csc_matrix((N, N), dtype=complex)
    for m in taps:
        Hs = Hs + sp.csc_matrix((h0[m:, m], (n_idx[m:], n_idx[m:] - m)), shape=(N, N))
    ncol = 300
    t_c = time.perf_counter()
    Dslice = Q.conj().T @ (Hs @ G[:, :ncol])
    cpu_s = (time.perf_counter() - t_c) * Ksym / ncol
    got = ctx.transmission_matrix_entries("F", 0, np.arange(0, 40), np.arange(0, 40))
    worst = max(worst, float(np.max(np.abs(got - np.diag(Dslice)[:40])) / np.max(np.abs(Dslice))))
    bytes_hg = 16.0 * sum(ctx_hg_rows(fb, taps, N)) * B            # H*G rows written per batch (algorithmic: 16 B per element)
    out = {"metric": METRIC, "value": B * K / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "scaled bandwidth (BASELINE.json configs[4]): FBMC L=%d subcarriers x 30 symbols, fs=%.2f MHz, N=%d samples, "
                                  "K=%d symbols, VehicularA (T=%d taps) at 500 km/h; per realization: NewRealization + banded H applied to G + "
                                  "D = Q^H H G (%.1f GB) + diag(D)  [rows A-D of SURVEY.md 8a]; the MMSE / IC rows are not run at this size "
                                  "(R_Dij_hP alone would be P x K^2 x 16 B = 259 GB for P = 200 pilots)" % (L, fs / 1e6, N, Ksym, T, 16.0 * Ksym * Ksym / 1e9),
                      "realizations_per_step": B, "timing": "CUDA events on the library's stream around all K steps",
                      "l2": "per realization 2.5 GB of H*G planes and 1.3 GB of D vs 126 MB L2"},
           "e2e": {"value": B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(du.nbytes + pu.nbytes), "d2h_bytes_per_step": int(hd.numel() * 8),
                   "api": "chest_new_realization (host draws) + chest_transmission_matrix_batch (C ABI): Doppler / phase uniforms in, diag(D) out; D stays in HBM"},
           "gpu_launches": int(launches), "clocks": clk, "setup_s": setup_s,
           "roofline": {"kernel": "k_gemm_d (D = Q^H (H G): persistent mbarrier-ring FP64 DMMA GEMM over the tile pairs whose supports overlap)",
                        "bound": "tensor", "achieved": B * K * flops / (gd_ms * 1e-3) / 1e12, "peak": peak_dmma, "unit": "TFLOP/s",
                        "frac": B * K * flops / (gd_ms * 1e-3) / 1e12 / peak_dmma, "traffic": None,
                        "algorithmic_flops_per_realization": flops, "dense_model_flops_per_realization": 8.0 * Ksym * Ksym * N + 8.0 * T * N * Ksym,
                        "avg_launch_ms": gd_ms / K, "peak_source": "FP64 DMMA loop measured in this run (chest_fp64_peak)"},
           "roofline_k1": {"kernel": "k_apply_hg (banded, never-materialised H applied to the columns of G)", "bound": "hbm",
                           "achieved": bytes_hg * K / (hg_ms * 1e-3) / 1e9 if hg_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                           "frac": bytes_hg * K / (hg_ms * 1e-3) / 1e9 / hbm_peak if hg_ms > 0 else None, "peak_source": hbm_src, "ms_per_step": hg_ms / K},
           "parity_sample": "ok (max deviation %.1e of max|D| over sampled entries)" % worst if worst < 1e-9 else "MISMATCH %.2e" % worst,
           "cpu_baseline": {"value": 1.0 / cpu_s, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "Q' * (H * G) of DS.m:388-389 with sparse H and dense G, Q (NumPy/OpenBLAS + SciPy), %d of %d columns of D timed and scaled; "
                                      "channel synthesis not included" % (ncol, Ksym)}}
    print(json.dumps(out), flush=True)
    ctx.close()
    if worst >= 1e-9:
        raise SystemExit(3)


def ctx_hg_rows(fb, taps, N):
    """Rows of H*G per column: the column's support widened by the largest delay (what k_apply_hg has to write)."""
    Np, TS, Ksym, L = fb.Nr["SamplesPrototypeFilter"], fb.Implementation["TimeSpacing"], fb.Nr["MCSymbols"], fb.Nr["Subcarriers"]
    return [min(N, k * TS + Np + int(taps[-1])) - k * TS for k in range(Ksym) for _ in range(L)]


def _one_json_line_stdout():
    """Libraries (NCCL's version banner, ...) write to file descriptor 1 behind Python's back; the contract is ONE
    JSON line on stdout.  Point fd 1 at stderr for the whole run and keep the real stdout for that line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


if __name__ == "__main__":
    _one_json_line_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "sweep":
        run_sweep(a)
    elif a.workload == "sv":
        run_sv(a)
    elif a.workload == "scaled":
        run_scaled(a)
    else:
        run_b200(a)
