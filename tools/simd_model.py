"""Offline model of the megakernel's warp scheduling (no GPU needed): per-ray traversal step counts come from the
host emulation of the kernel source (tests/hostemu, emu_pixel_costs); a warp of 32 persistent lanes is then stepped
through the regenerate / node-phase / triangle-drain loop of render.cu with instruction weights taken from the ncu
SASS breakdown (profiles/README.md: node visit 294, triangle test 100, shading + ray set-up ~300, votes ~12 per
loop iteration). Output: warp instructions issued, average active lanes, the split by phase — compared with what
ncu measured — and the same for what-if policies. A model of SIMD efficiency only: latency, caches and pipes are
not in it.

    python tools/simd_model.py [--workload c3_sponza_scale] [--width 160 --height 96] [--spp 16] [--refill 12]
"""
import argparse, ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, _hostemu
pkg = importlib.import_module("sycl-ray-tracer_b200")

W_NODE, W_TRI, W_SHADE, W_CAMERA, W_LOOP = 294, 100, 260, 90, 12


def collect(workload, w, h, spp, depth):
    data, _, _, _, d0 = bench.build_scene_data(workload)
    depth = depth or d0
    emu = _hostemu.Scene(data)
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    L = _hostemu.lib()
    L.emu_pixel_costs.restype = C.c_uint32
    L.emu_pixel_costs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    p = emu.cap.rt_render_params()
    p.max_depth, p.sample_count = depth, spp
    cap = spp * depth
    nn, nt, fs = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16), np.zeros(cap, np.uint8)
    pixels = []
    for by in range(0, h, 4):            # 8x4 block-major order, like the kernel
        for bx in range(0, w, 8):
            for i in range(32):
                x, y = bx + (i & 7), by + (i >> 3)
                if x >= w or y >= h:
                    continue
                n = L.emu_pixel_costs(emu.h, 0, C.addressof(cam.c), C.addressof(p), x, y, nn.ctypes.data, nt.ctypes.data, fs.ctypes.data, cap)
                pixels.append((nn[:n].copy(), nt[:n].copy(), fs[:n].copy()))
    return pixels


def simulate(pixels, n_warps, refill=12, pooled_tris=False, coop_tail=False, tstack=8):
    """returns dict of issued warp-instructions by phase and lane-weighted sums"""
    nxt = [0]
    stats = {k: [0.0, 0.0] for k in ("node", "tri", "shade", "loop")}      # [warp instructions, lane-instructions]

    def fetch():
        i = nxt[0]
        nxt[0] += 1
        return pixels[i] if i < len(pixels) else None

    warps = [dict(pix=[None] * 32, ray=[0] * 32, node=[0] * 32, tri=[0] * 32, pend=[0] * 32, trav=[False] * 32, need=[True] * 32, dead=[False] * 32) for _ in range(n_warps)]
    live = list(range(n_warps))
    clock = [0.0] * n_warps            # warp instructions issued by each warp: its time line (equal issue rates)
    while live:
        # the warp that is furthest behind runs next, so pixels are handed out in time order like the global counter
        for wi in sorted(live, key=lambda i: clock[i])[:1]:
            wp = warps[wi]
            before = sum(v[0] for v in stats.values())
            # ---- regenerate: lanes whose ray finished shade it / start the next ray / fetch a pixel
            shading = 0
            for l in range(32):
                if wp["dead"][l] or wp["trav"][l]:
                    continue
                cost_lane = False
                if not wp["need"][l]:                      # a finished ray: shade it
                    cost_lane = True
                    wp["ray"][l] += 1
                while True:
                    px = wp["pix"][l]
                    if px is None or wp["ray"][l] >= len(px[0]):
                        px = fetch()
                        wp["pix"][l], wp["ray"][l] = px, 0
                        if px is None:
                            wp["dead"][l] = True
                            break
                        if len(px[0]) == 0:
                            continue
                    r = wp["ray"][l]
                    wp["node"][l], wp["tri"][l], wp["pend"][l] = int(px[0][r]), int(px[1][r]), 0
                    wp["trav"][l], wp["need"][l] = True, False
                    cost_lane = True
                    break
                shading += cost_lane
            if shading:
                stats["shade"][0] += W_SHADE
                stats["shade"][1] += W_SHADE * shading
            if all(wp["dead"]):
                live.remove(wi)
                clock[wi] += sum(v[0] for v in stats.values()) - before
                continue
            # ---- traverse phase
            trav0 = [l for l in range(32) if wp["trav"][l]]
            thr = max(1, (len(trav0) * refill + 31) >> 5)
            while True:
                act = [l for l in trav0 if wp["node"][l] > 0]
                if not act:
                    break
                stats["node"][0] += W_NODE
                stats["node"][1] += W_NODE * len(act)
                stats["loop"][0] += W_LOOP
                stats["loop"][1] += W_LOOP * 32
                full = False
                for l in act:
                    # the ray's triangles surface evenly over its node steps
                    rem_nodes = wp["node"][l]
                    give = (wp["tri"][l] + rem_nodes - 1) // rem_nodes if wp["tri"][l] else 0
                    wp["node"][l] -= 1
                    wp["tri"][l] -= give
                    wp["pend"][l] += give
                    full |= wp["pend"][l] >= tstack * 2
                dry = sum(1 for l in trav0 if wp["node"][l] == 0)
                if dry >= thr or full:
                    break
            # ---- drain
            while True:
                must = [l for l in trav0 if wp["pend"][l] > 0 and (wp["node"][l] == 0 or wp["pend"][l] >= (tstack - 2) * 2)]
                if not must:
                    break
                have = [l for l in trav0 if wp["pend"][l] > 0]
                if pooled_tris:
                    total = sum(wp["pend"][l] for l in have)
                    rounds = -(-total // 32)
                    stats["tri"][0] += rounds * (W_TRI + 35)
                    stats["tri"][1] += (W_TRI + 35) * total
                    for l in have:
                        wp["pend"][l] = 0
                else:
                    stats["tri"][0] += W_TRI
                    stats["tri"][1] += W_TRI * len(have)
                    for l in have:
                        wp["pend"][l] -= 1
            for l in trav0:
                if wp["node"][l] == 0 and wp["pend"][l] == 0:
                    wp["trav"][l] = False
            clock[wi] += sum(v[0] for v in stats.values()) - before
    stats["_makespan"] = [max(clock), sum(clock) / n_warps]
    return stats


def simulate_k(pixels, n_warps, refill=12, K=2, tstack=8):
    """what-if: every lane owns K independent pixels (ray contexts) and advances whichever has node work — the
    software-pipelined "two rays per lane" idea; costs registers on the GPU, which the model does not see"""
    nxt = [0]
    stats = {k: [0.0, 0.0] for k in ("node", "tri", "shade", "loop")}

    def fetch():
        i = nxt[0]
        nxt[0] += 1
        return pixels[i] if i < len(pixels) else None

    def new_ctx():
        return dict(pix=None, ray=0, node=0, tri=0, pend=0, trav=False, need=True, dead=False)

    warps = [[[new_ctx() for _ in range(K)] for _ in range(32)] for _ in range(n_warps)]
    live = list(range(n_warps))
    while live:
        for wi in list(live):
            wp = warps[wi]
            rounds, lane_work = 0, 0
            for l in range(32):
                mine = 0
                for c in wp[l]:
                    if c["dead"] or c["trav"]:
                        continue
                    did = False
                    if not c["need"]:
                        did = True
                        c["ray"] += 1
                    while True:
                        px = c["pix"]
                        if px is None or c["ray"] >= len(px[0]):
                            px = fetch()
                            c["pix"], c["ray"] = px, 0
                            if px is None:
                                c["dead"] = True
                                break
                            if len(px[0]) == 0:
                                continue
                        r = c["ray"]
                        c["node"], c["tri"], c["pend"], c["trav"], c["need"] = int(px[0][r]), int(px[1][r]), 0, True, False
                        did = True
                        break
                    mine += did
                rounds = max(rounds, mine)
                lane_work += mine
            if rounds:
                stats["shade"][0] += W_SHADE * rounds
                stats["shade"][1] += W_SHADE * lane_work
            if all(c["dead"] for l in range(32) for c in wp[l]):
                live.remove(wi)
                continue
            ctxs = [(l, c) for l in range(32) for c in wp[l] if c["trav"]]
            n_trav_lanes = len({l for l, _ in ctxs})
            thr = max(1, (n_trav_lanes * refill + 31) >> 5)
            while True:
                pick = {}
                for l, c in ctxs:
                    if c["node"] > 0 and l not in pick:
                        pick[l] = c
                if not pick:
                    break
                stats["node"][0] += W_NODE
                stats["node"][1] += W_NODE * len(pick)
                stats["loop"][0] += W_LOOP
                stats["loop"][1] += W_LOOP * 32
                full = False
                for c in pick.values():
                    give = (c["tri"] + c["node"] - 1) // c["node"] if c["tri"] else 0
                    c["node"] -= 1
                    c["tri"] -= give
                    c["pend"] += give
                    full |= c["pend"] >= tstack * 2
                # a lane is "dry" when it has a finished context waiting to be shaded and no node work left in the other
                finished = sum(1 for l, c in ctxs if c["node"] == 0 and c["pend"] == 0)
                idle_lanes = 32 - len({l for l, c in ctxs if c["node"] > 0})
                if finished >= thr * K or idle_lanes >= thr or full:
                    break
            while True:
                must = [c for l, c in ctxs if c["pend"] > 0 and (c["node"] == 0 or c["pend"] >= (tstack - 2) * 2)]
                if not must:
                    break
                pick = {}
                for l, c in ctxs:
                    if c["pend"] > 0 and l not in pick:
                        pick[l] = c
                stats["tri"][0] += W_TRI
                stats["tri"][1] += W_TRI * len(pick)
                for c in pick.values():
                    c["pend"] -= 1
            for l, c in ctxs:
                if c["node"] == 0 and c["pend"] == 0:
                    c["trav"] = False
    return stats


def report(tag, st):
    span = st.pop("_makespan", None)
    tot_w = sum(v[0] for v in st.values())
    tot_l = sum(v[1] for v in st.values())
    parts = "  ".join(f"{k} {100 * v[0] / tot_w:4.1f}% @ {v[1] / max(v[0], 1):4.1f}" for k, v in st.items())
    tail = f"   makespan / mean warp time {span[0] / span[1]:.3f}" if span else ""
    print(f"{tag:34s} warp-instr {tot_w / 1e6:8.2f} M   avg lanes {tot_l / tot_w:5.2f}   {parts}{tail}")
    return tot_w


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_sponza_scale")
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--height", type=int, default=96)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--warps", type=int, default=48)
    a = ap.parse_args()
    px = collect(a.workload, a.width, a.height, a.spp, a.depth)
    rays = sum(len(p[0]) for p in px)
    print(f"{a.workload}: {len(px)} pixels, {rays} rays, {sum(int(p[0].sum()) for p in px) / rays:.2f} node steps/ray, {sum(int(p[1].sum()) for p in px) / rays:.2f} tri tests/ray")
    base = report("current policy (refill 12)", simulate(px, a.warps, 12))
    for r in (4, 8, 16, 24):
        report(f"refill {r}", simulate(px, a.warps, r))
    t = report("pooled triangle drain", simulate(px, a.warps, 12, pooled_tris=True))
    print(f"pooled drain: {100 * (base - t) / base:.1f} % fewer warp instructions")
    # the tail: few pixels per lane (as under 8-way tile sharding) and the order they are handed out in
    cost = [float(p[0].sum()) * W_NODE + float(p[1].sum()) * W_TRI + len(p[0]) * W_SHADE for p in px]
    blocks = [list(range(i, min(i + 32, len(px)))) for i in range(0, len(px), 32)]
    def by_blocks(order):
        return [px[i] for b in order for i in blocks[b]]
    bcost = [sum(cost[i] for i in b) for b in blocks]
    many = max(2, len(px) // (32 * 5))                           # ~5 pixels per lane
    report(f"tail: image order, {many} warps", simulate(px, many, 12))
    report("tail: expensive blocks first", simulate(by_blocks(sorted(range(len(blocks)), key=lambda b: -bcost[b])), many, 12))
    report("tail: cheap blocks first", simulate(by_blocks(sorted(range(len(blocks)), key=lambda b: bcost[b])), many, 12))
    for K in (2, 3):
        t = report(f"{K} ray contexts per lane", simulate_k(px, max(1, a.warps // K), 12, K))
        print(f"{K} contexts: {100 * (base - t) / base:.1f} % fewer warp instructions")
