"""Per-kernel SASS evidence for profiles/: registers, spill bytes, shared memory (cuobjdump --dump-resource-usage) and
the counts of the opcodes that say what a kernel is made of -- fp64 / fp32 arithmetic, tensor-memory traffic
(LDTM / STTM / UTCATOMSWS = tcgen05.ld / .st / .alloc), TMA (UTMALDG / UTMASTG / UTMAPF = cp.async.bulk.prefetch.tensor / UBLKCP), tensor-core MMA (UTC*MMA), local
memory (LDL / STL = spills), shared and global memory instructions.
    python scripts/sass_summary.py [path/to/libcaf_b200.so] > profiles/r02_sass_summary.txt
Needs no GPU."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "caf_cookoff_b200", "libcaf_b200.so")
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1); continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = collections.OrderedDict()
cur = None
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m: arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
cols = ["DFMA", "DADD", "DMUL", "FFMA", "FFMA2", "FADD2", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCHMMA",
        "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "SYNCS", "SHFL"]
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print(f"# SASS summary of {os.path.relpath(so, ROOT)}: arch = {sorted(arch)}, {len(kern)} kernels (static instruction counts)")
print("# REG = registers per thread, STACK = stack frame bytes (spill stores live there), SMEM = static shared bytes; opcode columns are static counts")
tot = collections.Counter()
for name, c in kern.items():
    reg, stack, shared, local = usage.get(name, (0, 0, 0, 0))
    d = demangle(name)
    d = re.sub(r"caf::", "", d)
    d = re.sub(r"\(.*", "", d)
    n = sum(c.values())
    parts = [f"{k}={c[k]}" for k in cols if c[k]]
    any_mma = sum(v for k, v in c.items() if "MMA" in k)
    print(f"{d}\n    REG={reg} STACK={stack} SMEM={shared} LOCAL={local} INSTR={n}  " + " ".join(parts) + (f" MMA={any_mma}" if any_mma else ""))
    tot.update(c)
print("# totals over all kernels: " + " ".join(f"{k}={tot[k]}" for k in cols) + f" any-MMA={sum(v for k, v in tot.items() if 'MMA' in k)}")
