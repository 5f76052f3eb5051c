"""Attributes ncu warp-stall samples of one kernel to CUDA source lines.
usage: python scripts/ncu_lines.py <rep.ncu-rep> <kernel substring> [top N]
Joins `ncu --page source --csv` (SASS view with per-instruction samples) with `nvdisasm -g` line markers of the in-tree .so."""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
# one section per profiled launch: ["Kernel Name", name] / header / data...; take the first section whose kernel matches `pat`
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
import re as _re
want = _re.sub(r"I?L[ib](\d+)E", r"\1", pat)
sec = next((k for k in range(len(starts) - 1) if all(tok in allrows[starts[k]][1].replace("(int)", "").replace(" ", "") for tok in _re.findall(r"[A-Za-z_0-9]+", want) if not tok.isdigit() or True)), 0)
rows = allrows[starts[sec]:starts[sec + 1]]
print("section:", rows[0][1][:90])
hdr = rows[1]
iA, iS, iSamp, iInst = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][iA], 16)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "fem_2d_b200", "libfem2d_b200.so")], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if pat not in txt:
        continue
    cur_fn, cur_line, infn = None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = pat in m.group(1); continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of.setdefault(int(m.group(1), 16), cur_line)
    if line_of:
        break

agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in data:
    off = int(r[iA], 16) - base
    key = line_of.get(off, ("?", 0))
    s = int(r[iSamp] or 0); tot += s
    a = agg[key]; a[0] += s; a[1] += int(r[iInst] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: a[2][hdr[i]] += v
src_cache = {}
def src(key):
    f, l = key
    for d in ("fem_2d_b200/csrc", "include"):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:110] if 0 < l <= len(src_cache[p]) else ""
    return ""
print(f"total samples {tot}")
for key, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
    print(f"{100*s/max(tot,1):5.1f}%  inst={n:9d}  {key[0]}:{key[1]:4d}  [{top}]  {src(key)}")
