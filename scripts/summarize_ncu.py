"""Turn gpurun_out/*.ncu-rep / launch CSVs into the small text summaries committed under profiles/.
   python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv r01"""
import collections, csv, io, json, os, subprocess, sys

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

def ncu_csv(page):
    txt = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))

raw = ncu_csv("raw")
hdr, units = raw[0], raw[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum"]
lines, summary = [], {}
for r in raw[2:]:
    name = r[hdr.index("Kernel Name")]
    lines.append(f"== {name}")
    for k in keys:
        if k in hdr:
            i = hdr.index(k); lines.append(f"   {k} = {r[i]} {units[i]}")
            summary[k] = (r[i], units[i])
    for i, k in enumerate(hdr):
        if "average_warps_issue_stalled" in k and "per_issue_active" in k:
            try:
                if float(r[i]) > 0.05: lines.append(f"   stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {r[i]} warps/issue")
            except ValueError: pass
def to_bytes(v, u):
    f = float(v); return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
traffic = None
if "dram__bytes_read.sum" in summary:
    traffic = to_bytes(*summary["dram__bytes_read.sum"]) + to_bytes(*summary["dram__bytes_write.sum"])
src = ncu_csv("source")
if len(src) > 2:
    h = src[1]; isrc, isamp, iexe = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    data = src[2:]; tot = collections.Counter()
    for r in data:
        for i, x in stalls: tot[x] += int(r[i] or 0)
    lines.append("== warp-state samples over the whole kernel: " + ", ".join(f"{k[6:]} {v}" for k, v in tot.most_common(9)))
    # row loop vs everything else: an instruction of the loop runs once per (row, warp) -- 400 rows x 8 or 16 warps --
    # while set-up / prologue / tail code runs once per warp and CTA
    import re
    def fl(x):
        try: return float(x)
        except ValueError: return 0.0
    loop = [r for r in data if fl(r[iexe]) >= 3000]
    rest = [r for r in data if fl(r[iexe]) < 3000]
    for name, part in (("row loop", loop), ("set-up, prologue, H exchange, tail", rest)):
        c = collections.Counter()
        for r in part:
            for i, x in stalls: c[x] += int(r[i] or 0)
        n = sum(int(r[isamp] or 0) for r in part); t = sum(c.values()) or 1
        lines.append(f"== {name}: {n} samples ({100.0 * n / max(1, sum(int(r[isamp] or 0) for r in data)):.0f} % of the kernel): "
                     + ", ".join(f"{k[6:]} {100.0 * v / t:.0f}%" for k, v in c.most_common(9)))
    ops = collections.Counter()
    for r in loop:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
        if m: ops[m.group(2)] += fl(r[iexe])
    lines.append("== row loop, instructions per warp and row (executions / 6400): "
                 + ", ".join(f"{k} {v / 6400:.0f}" for k, v in ops.most_common(12)))
    lines.append("== top instructions by samples (samples, executions, SASS, top stall)")
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:15]:
        st = sorted(((x, int(r[i] or 0)) for i, x in stalls), key=lambda t: -t[1])[0]
        lines.append(f"   {r[isamp]:>5} {r[iexe]:>7}  {r[isrc][:64]:64s} {st[0]}={st[1]}")
open(os.path.join(out_dir, f"{tag}_ncu_full_rows_kernel.txt"), "w").write("\n".join(lines) + "\n")
json.dump({"kernel": "caf_rows_kernel<double, kSurface, FULL>", "dram_bytes_per_launch": traffic,
           "source": os.path.basename(rep), "note": "ncu --set full --clock-control none, one launch"},
          open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)
# launch list: keep kernel name + duration only
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
h = rows[0]; ik, iv, ig = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("id,kernel,grid,gpu__time_duration_ns\n")
    agg = collections.defaultdict(list)
    for n, r in enumerate(rows[1:]):
        k = r[ik].split("(")[0][:90]
        f.write(f"{n},{k},{r[ig]},{r[iv]}\n"); agg[k].append(float(r[iv].replace(',', '')))
    f.write("# per-kernel share of the profiled region (cold-cache, serialised launches: compare shares, not absolutes)\n")
    total = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda t: -sum(t[1])):
        f.write(f"# {k}: launches {len(v)}, mean {sum(v)/len(v):.0f} ns, share {100*sum(v)/total:.1f} %\n")
print("\n".join(lines[:40])); print("traffic bytes/launch:", traffic)
