"""Summarise an .ncu-rep (run in the build container): key raw metrics, stall mix, and per-phase
(barrier-delimited) sample shares from the source page.  Usage: ncu_summary.py <rep> [out.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "sm__icc_requests.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active"]
k = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
print("kernel:", vals[k] if k is not None else "?", file=out)
if len(rows) > 3:                                  # several launches captured: one line each, details below are the first one's
    cols = [hdr.index(h) for h in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum") if h in hdr]
    for i, r in enumerate(rows[2:]):
        if len(r) > max(cols):
            print(f"launch {i}: " + ", ".join(f"{hdr[c]} = {r[c]} {units[c]}" for c in cols), file=out)
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        print(f"{h:85s} {u:16s} {v}", file=out)
st = []
for h, u, v in zip(hdr, units, vals):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st.append((float(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(x for x, _ in st) or 1
print("stall mix: " + ", ".join(f"{h} {100*x/tot:.1f}%" for x, h in sorted(st, reverse=True)[:9]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, h):
    try:
        return float(r[ix[h]])
    except (ValueError, IndexError, KeyError):
        return 0.0
def op(r):
    t = r[ix["Source"]].split()
    if not t: return ""
    return (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
# a report with several launches repeats the source page once per launch: keep the first copy
if "Address" in ix:
    first = data[0][ix["Address"]] if data and len(data[0]) > ix["Address"] else None
    for i in range(1, len(data)):
        if len(data[i]) > ix["Address"] and data[i][ix["Address"]] == first:
            data = data[:i]
            break
segs, start = [], 0
for i, r in enumerate(data):
    if op(r) == "BAR":
        segs.append((start, i)); start = i + 1
segs.append((start, len(data)))
tot = sum(f(r, "# Samples") for r in data) or 1
keys = ["stall_wait", "stall_selected", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_math", "stall_mio", "stall_barrier"]
print(f"SASS instructions: {len(data)} ({len(data)*16/1024:.0f} KiB)", file=out)
print("phase(barrier-delimited)  instrs samples% exec(M) dp%  | " + " ".join(x[6:13] for x in keys), file=out)
for a, b in segs:
    rs = data[a:b + 1]
    s = sum(f(r, "# Samples") for r in rs)
    ex = sum(f(r, "Instructions Executed") for r in rs)
    dp = sum(f(r, "Instructions Executed") for r in rs if op(r) in ("DADD", "DFMA", "DMUL"))
    if s / tot < 0.002: continue
    print(f"{a:5d}-{b:5d} {b-a+1:6d} {100*s/tot:7.2f} {ex/1e6:9.1f} {100*dp/max(ex,1):5.1f} | " +
          " ".join(f"{100*sum(f(r,k) for r in rs)/max(s,1):7.1f}" for k in keys), file=out)
