import csv, subprocess, sys, collections, re, os, tempfile
rep, kname, so = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = []
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname not in txt: continue
    infn, cur = False, ("?", 0)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
        if m: infn = kname in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m: lines.append((int(m.group(1), 16), cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi, end = his[0], (his[1] - 2 if len(his) > 1 else len(rows))
hdr, data = rows[hi], [r for r in rows[hi + 1:end] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(lines)
tot = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
# contiguous regions of similar execution count
reg = []
for r, (addr, loc, txt) in zip(data, lines):
    ie = float(r[ix["Instructions Executed"]] or 0); te = float(r[ix["Thread Instructions Executed"]] or 0)
    reg.append((addr, ie, te, loc, txt))
# group into chunks of 64 instructions
CH = int(sys.argv[4]) if len(sys.argv) > 4 else 64
for i in range(0, len(reg), CH):
    c = reg[i:i + CH]
    ie = sum(x[1] for x in c); te = sum(x[2] for x in c)
    if ie / tot < 0.002: continue
    locs = collections.Counter(x[3] for x in c).most_common(2)
    print(f"{c[0][0]:6x}-{c[-1][0]:6x}  inst {ie / tot * 100:5.2f}%  lanes {te / max(ie, 1):5.1f}  exec/instr {ie / len(c) / 1e6:8.1f}M   {locs[0][0][0]}:{locs[0][0][1]} {locs[1][0][0] if len(locs) > 1 else ''}:{locs[1][0][1] if len(locs) > 1 else ''}")
