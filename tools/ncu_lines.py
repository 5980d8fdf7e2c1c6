"""Join an ncu SASS source page with nvdisasm line info: per source line (innermost inlined frame)
share of warp instructions, stall samples and average active threads.
usage: ncu_lines.py report.ncu-rep kernel_substring [cubin]"""
import csv, collections, os, re, subprocess, sys, tempfile
rep, kname = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("RT_LIB_PATH") or os.path.join(root, "sycl-ray-tracer_b200", "librt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = []
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname not in txt:
        continue
    infn, cur = False, ("?", 0)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
        if m:
            infn = kname in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi, end = his[0], (his[1] - 2 if len(his) > 1 else len(rows))  # first kernel of the report
hdr, data = rows[hi], [r for r in rows[hi + 1:end] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r, (_, loc, _) in zip(data, lines):
    a = agg[loc]
    a[0] += float(r[ix["Instructions Executed"]] or 0)
    a[1] += float(r[ix["Thread Instructions Executed"]] or 0)
    a[2] += float(r[ix["# Samples"]] or 0)
tot = sum(a[0] for a in agg.values()); smp = sum(a[2] for a in agg.values())
srcs = {}
def text(loc):
    f, l = loc
    if f not in srcs:
        p = os.path.join(root, "sycl-ray-tracer_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][l - 1].strip()[:80] if 0 < l <= len(srcs[f]) else ""
print(f"{'inst%':>6} {'smp%':>6} {'thr':>5}  location")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{a[0]/tot*100:6.2f} {a[2]/smp*100:6.2f} {a[1]/max(a[0],1):5.1f}  {loc[0]}:{loc[1]}  {text(loc)}")
