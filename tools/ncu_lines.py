"""Summarise an ncu report: headline metrics + stall samples aggregated per CUDA source line."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sel = ["--kernel-id", kid] if kid else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}] = {[r[i] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, hdr, agg = None, None, collections.OrderedDict()
stall_tot = collections.Counter()
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or not r:
        continue
    off = len(r) - len(hdr)
    if r[0].isdigit():
        line = (cur, int(r[0]), ",".join(r[1:2 + off]).strip()[:80])
        continue
    if r[0] == "" and off >= 0 and len(r) > 6 and r[2 + 0].startswith("0x"):
        try:
            s = int(r[hdr.index("Warp Stall Sampling (All Samples)") + off])
            ie = int(r[hdr.index("Instructions Executed") + off])
        except (ValueError, IndexError):
            continue
        key = r[2 + off] if off else r[2]
        agg.setdefault(key, [0, 0, None, None])
        # each SASS address may be listed under several inlined lines: keep the first listing only
        if agg[key][2] is None:
            agg[key] = [s, ie, line, r[3 + off if off else 3].strip()[:60]]
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        stall_tot[h] += int(r[hdr.index(h) + off] or 0)
                    except ValueError:
                        pass
per_line = collections.Counter()
ie_line = collections.Counter()
for s, ie, line, sass in agg.values():
    per_line[line] += s
    ie_line[line] += ie
S = sum(per_line.values())
print(f"--- total samples {S}, warp instructions {sum(ie_line.values())}")
for k, v in stall_tot.most_common(8):
    print(f"   {k:26s}{100 * v / max(S, 1):6.1f}%")
print("--- top source lines by stall samples")
for line, s in per_line.most_common(top):
    print(f"{100 * s / S:5.1f}%  ie={ie_line[line]:11d}  {line[0]}:{line[1]}  {line[2]}")
