"""Development tool: key metrics, stall reasons and hot instructions of one kernel from an .ncu-rep.
usage: python tools/ncu_report.py file.ncu-rep [n_hot] [kernel index in the report]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2 + kidx]
d = {h[i]: (v[i], u[i]) for i in range(len(h))}
print(d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size"))
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]:
    print("  %-75s %s" % (k, d.get(k)))
st = [(k, float(d[k][0].replace(",", ""))) for k in d if k.startswith("smsp__average_warp") and "issue_stalled" in k and "ratio" in k]
print("stalls per issue:", ", ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), x)
                                      for k, x in sorted(st, key=lambda t: -t[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(kidx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
data = [x for x in rows[2:] if len(x) == len(hdr)]
def f(x, k):
    try: return float(x[ix[k]] or 0)
    except Exception: return 0.0
tot = sum(f(x, "# Samples") for x in data)
print("samples", tot)
for x in sorted(data, key=lambda x: -f(x, "# Samples"))[:n_hot]:
    stalls = {k: f(x, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    big = sorted(stalls.items(), key=lambda t: -t[1])[:3]
    print("%5.1f%%  %-64s %s" % (100 * f(x, "# Samples") / tot, x[ix["Source"]][:64], " ".join("%s=%d" % (k[6:], n) for k, n in big)))
print("shared wavefronts: total %.3g" % sum(f(x, "L1 Wavefronts Shared") for x in data))
for x in sorted(data, key=lambda x: -f(x, "L1 Wavefronts Shared"))[:6]:
    print("   %-60s total %.3g ideal %.3g exec %s" % (x[ix["Source"]][:60], f(x, "L1 Wavefronts Shared"), f(x, "L1 Wavefronts Shared Ideal"), x[ix["Instructions Executed"]]))
