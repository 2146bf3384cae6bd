"""Summary of one `ncu --set full` capture for profiles/: key raw metrics + hot-loop share + per-file attribution.
   python tools/ncu_summary.py rep.ncu-rep "<title>" > profiles/ncu_<round>_<what>_summary.txt"""
import csv, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(raw.split("\n")))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
print(title)
print("ncu --set full --clock-control none (serialised, cold caches: shares, not absolute times, are what carries over)")
for i, h in enumerate(hdr):
    if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print("%s | %s | %s" % (h, units[i], vals[i]))
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(src.split("\n")))
hi = 0 if 'Address' in rows[0] else 1
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) > 10]
S = lambda r: int(r[ix['# Samples']] or 0)
N = lambda r: int(r[ix['Instructions Executed']] or 0)
tot_s = sum(map(S, data)); tot_i = sum(map(N, data)); nmax = max(map(N, data))
hot = [r for r in data if N(r) > 0.6 * nmax]
n_ffma2 = sum(1 for r in hot if "FFMA2" in r[ix['Source']])
print("--- hottest loop (instructions executed > 0.6 x max): %d SASS instructions per trip (%d FFMA2), %.4g trips, %.1f%% of warp samples, %.1f%% of executed instructions"
      % (len(hot), n_ffma2, nmax, 100.0 * sum(map(S, hot)) / tot_s, 100.0 * sum(map(N, hot)) / tot_i))
stalls = [h for h in hdr if h.startswith('stall_') and "Not Issued" not in h]
agg = {}
for r in data:
    for h in stalls:
        v = int(r[ix[h]] or 0)
        if v: agg[h[6:]] = agg.get(h[6:], 0) + v
print("--- warp-sample stall mix (whole kernel):", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
mix = {}
for r in data:
    op = r[ix['Source']].split()
    op = [t for t in op if not t.startswith("@")][0].split(".")[0] if op else "?"
    mix[op] = mix.get(op, 0) + N(r)
print("--- executed instruction mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
