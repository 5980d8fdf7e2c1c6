"""profiles/roofline_traffic.json from ncu captures: python tools/update_traffic.py <tag> <workload> <renderer> <report.ncu-rep> <plain.log> <summary.txt>
(dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, per ray of that launch: bench.py scales it to its own rays)"""
import csv, json, os, re, subprocess, sys
tag, wl, rend, rep, plain, summ = sys.argv[1:7]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name)
    v, u = float(r[i]), units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
ms = float(r[hdr.index("gpu__time_duration.sum")]) * {"ms": 1, "us": 1e-3, "msecond": 1, "usecond": 1e-3, "second": 1e3, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1)
line = [l for l in open(plain) if " rays, " in l][-1]
rays = int(re.search(r"(\d+) rays", line).group(1))
m = re.search(r"(\d+x\d+) spp=(\d+)", line)
p = os.path.join(root, "profiles", "roofline_traffic.json")
d = json.load(open(p))
e = d.setdefault(wl, {}).setdefault(rend, {})
e.update({"dram_bytes_per_ray": round(dram / rays, 3), "dram_bytes_per_launch_of_capture": int(dram), "capture": summ,
          "launch": f"{'k_megakernel' if rend == 'megakernel' else 'k_wf_flow'}, {m.group(1)}, {m.group(2)} spp, depth 10 ({ms:.1f} ms)",
          "l1_hit_pct": round(float(r[hdr.index('l1tex__t_sector_hit_rate.pct')]), 1), "l2_hit_pct": round(float(r[hdr.index('lts__t_sector_hit_rate.pct')]), 1)})
json.dump(d, open(p, "w"), indent=1)
print(tag, wl, rend, f"{dram / rays:.2f} DRAM bytes per ray, {dram / 1e9:.3f} GB per launch, {ms:.1f} ms")
