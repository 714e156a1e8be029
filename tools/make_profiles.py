"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- launch list: per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("<")[0].replace("void ", "")
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += float(r[-1]) / 1e6
total_ms = sum(v[1] for v in tot.values())
with open(os.path.join(P, f"launches_{tag}.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `python bench.py --steps 2 --warmup 1 --skip-cpu --fields 32 --gs-events 1024`, N=1\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised launches: "
            "compare SHARES, not absolute times; the first 400 launches of the command).\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / total_ms:.1f}% |\n")
    f.write(f"\nTotal {total_ms:.1f} ms over {len(rows)} launches.\n")
import shutil
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"launches_{tag}.csv"))

# ---- full captures -> key metrics
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
roof = {}
plain = json.load(open(os.path.join(G, "bench_plain.json")))
units_per_launch = {"sweep_bricks16_kernel": plain["config"]["fields_per_gpu"] * plain["config"]["grid"][0] ** 3 * 8,
                    "locate_uniform_kernel": int(plain["events"]["config"]["workload"].split(": ")[1].split(" events")[0])}
out = open(os.path.join(P, f"ncu_summary_{tag}.md"), "w")
out.write(f"# ncu --set full summaries ({tag})\n\nCaptured with `tools/profile_round.sh` (clock control none, one launch each, "
          "from the bench command).  Numbers under ncu are never bench values.\n")
for rep, kern in (("prof_fsm.ncu-rep", "sweep_bricks16_kernel"), ("prof_gs.ncu-rep", "locate_uniform_kernel")):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, vals = r[0], r[1], r[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    out.write(f"\n## `{m['Kernel Name'][1]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k in want:
        if k in m:
            out.write(f"| {k} | {m[k][1]} | {m[k][0]} |\n")
    def tobytes(key):
        u, v = m[key]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    roof[kern] = {"dram_bytes_per_launch": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
                  "dram_bytes_read": tobytes("dram__bytes_read.sum"), "dram_bytes_write": tobytes("dram__bytes_write.sum"),
                  "duration_ms_under_ncu": float(m["gpu__time_duration.sum"][1]) * (1.0 if m["gpu__time_duration.sum"][0] == "ms" else 1e-3),
                  "units_per_launch": units_per_launch[kern],
                  "unit": "node-updates" if kern.startswith("sweep") else "events",
                  "source": f"profiles/ncu_summary_{tag}.md ({rep}); command: python bench.py --steps 2 --warmup 1 --skip-cpu --fields 32 --gs-events 1024"}
    roof[kern]["dram_bytes_per_unit"] = roof[kern]["dram_bytes_per_launch"] / units_per_launch[kern]
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), path, "", "14"],
                           capture_output=True, text=True).stdout
    out.write("\nStall reasons and hottest source lines (warp-state sampling):\n\n```\n" + lines[lines.index("--- total samples"):] + "```\n")
# ---- instruction mix of the sweep kernel: warp instructions per node-update and pipe
pp = os.path.join(G, "pipes_fsm.csv")
if os.path.exists(pp):
    rows = [r for r in csv.reader(open(pp)) if len(r) > 10 and r[0].isdigit()]
    upd = units_per_launch["sweep_bricks16_kernel"]
    out.write(f"\n## Instruction mix of `sweep_bricks16_kernel` (one launch, {upd / 1e9:.2f} G node-updates)\n\n"
              "Warp-level instructions executed per pipe (`smsp__inst_executed_pipe_*.sum`), per 64 node-updates (= one warp-step of the "
              "brick walk) and per node-update (x 32 lanes / 64).  A pipe's share need not add up: an instruction can use two pipes.\n\n"
              "| metric | total | per 64 updates | lane-instructions per update |\n|---|---:|---:|---:|\n")
    for r in rows:
        name, val = r[-3], float(r[-1].replace(",", ""))
        out.write(f"| {name} | {val:.4g} | {val / upd * 64:.1f} | {val / upd * 32:.1f} |\n")
out.close()
json.dump(roof, open(os.path.join(P, "roofline.json"), "w"), indent=1)
print(open(os.path.join(P, f"launches_{tag}.md")).read())
print(json.dumps(roof, indent=1))
