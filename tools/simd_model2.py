"""Offline model of the round-2 megakernel scheduling (no GPU): K parked ray contexts per lane, switched only at
ray boundaries.

A lane owns K pixels ("contexts"). One context at a time is under traversal (its state lives in registers); the
others are parked in local memory as either READY (a ray waiting to be traced) or HIT (a closest hit waiting to be
shaded). The warp loops over four warp-uniform phases:
    shade : every lane with a HIT context shades ONE of them (-> READY, or the lane's pixel ends and a new one is fetched);
            run when >= t_shade lanes have a HIT context or >= t_idle lanes can neither traverse nor start
    start : lanes without a ray under traversal pick a READY context and initialise its traversal
    node  : node steps until `refill` of the traversing lanes ran dry (or a triangle stack is full)
    tri   : batched triangle drain
Per-ray step counts come from the host emulation of the kernel source (tools/simd_model.py collect()).

    python tools/simd_model2.py [--workload c3_sponza_scale] [--width 160 --height 96] [--spp 8]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import simd_model as m1

W_NODE, W_TRI, W_SHADE, W_START, W_LOOP = 250, 105, 190, 100, 12
READY, TRAV, HIT, DEAD = 0, 1, 2, 3


def simulate(pixels, n_warps, K=2, refill=8, t_shade=20, t_idle=4, tstack=8, pooled=False):
    nxt = [0]
    stats = {k: [0.0, 0.0] for k in ("node", "tri", "shade", "start", "loop")}

    def fetch():
        i = nxt[0]
        nxt[0] += 1
        return pixels[i] if i < len(pixels) else None

    def next_ray(c):
        """advance context c to its next ray (after shading); False when no pixel is left"""
        while True:
            px = c["pix"]
            if px is None or c["ray"] >= len(px[0]):
                px = fetch()
                c["pix"], c["ray"] = px, 0
                if px is None:
                    c["st"] = DEAD
                    return False
                if len(px[0]) == 0:
                    continue
            c["st"] = READY
            return True

    warps = []
    for _ in range(n_warps):
        wp = []
        for l in range(32):
            ctxs = [dict(pix=None, ray=0, st=HIT, node=0, tri=0, pend=0, first=True) for _ in range(K)]
            wp.append(dict(ctx=ctxs, cur=None))
        warps.append(wp)
    live = list(range(n_warps))
    clock = [0.0] * n_warps
    while live:
        wi = min(live, key=lambda i: clock[i])
        wp = warps[wi]
        before = sum(v[0] for v in stats.values())
        # ---- shade passes
        while True:
            pending = [l for l in range(32) if any(c["st"] == HIT for c in wp[l]["ctx"])]
            can_go = [l for l in range(32) if wp[l]["cur"] is not None or any(c["st"] == READY for c in wp[l]["ctx"])]
            idle = [l for l in pending if l not in can_go]
            if not pending:
                break
            if not (len(pending) >= t_shade or len(idle) >= t_idle or not can_go):
                break
            stats["shade"][0] += W_SHADE
            stats["shade"][1] += W_SHADE * len(pending)
            for l in pending:
                c = next(c for c in wp[l]["ctx"] if c["st"] == HIT)
                if not c["first"]:
                    c["ray"] += 1
                c["first"] = False
                next_ray(c)
        # ---- start
        starters = [l for l in range(32) if wp[l]["cur"] is None and any(c["st"] == READY for c in wp[l]["ctx"])]
        if starters:
            stats["start"][0] += W_START
            stats["start"][1] += W_START * len(starters)
            for l in starters:
                c = next(c for c in wp[l]["ctx"] if c["st"] == READY)
                px, r = c["pix"], c["ray"]
                c["node"], c["tri"], c["pend"], c["st"] = int(px[0][r]), int(px[1][r]), 0, TRAV
                wp[l]["cur"] = c
        trav0 = [l for l in range(32) if wp[l]["cur"] is not None]
        if not trav0:
            if all(c["st"] == DEAD for l in range(32) for c in wp[l]["ctx"]):
                live.remove(wi)
            clock[wi] += sum(v[0] for v in stats.values()) - before
            continue
        thr = max(1, (len(trav0) * refill + 31) >> 5)
        while True:
            act = [l for l in trav0 if wp[l]["cur"]["node"] > 0]
            if not act:
                break
            stats["node"][0] += W_NODE
            stats["node"][1] += W_NODE * len(act)
            stats["loop"][0] += W_LOOP
            stats["loop"][1] += W_LOOP * 32
            full = False
            for l in act:
                c = wp[l]["cur"]
                give = (c["tri"] + c["node"] - 1) // c["node"] if c["tri"] else 0
                c["node"] -= 1
                c["tri"] -= give
                c["pend"] += give
                full |= c["pend"] >= tstack * 2
            dry = sum(1 for l in trav0 if wp[l]["cur"]["node"] == 0)
            if dry >= thr or full:
                break
        while True:
            must = [l for l in trav0 if wp[l]["cur"]["pend"] > 0 and (wp[l]["cur"]["node"] == 0 or wp[l]["cur"]["pend"] >= (tstack - 2) * 2)]
            if not must:
                break
            have = [l for l in trav0 if wp[l]["cur"]["pend"] > 0]
            if pooled:
                total = sum(wp[l]["cur"]["pend"] for l in have)
                rounds = -(-total // 32)
                stats["tri"][0] += rounds * (W_TRI + 45)
                stats["tri"][1] += (W_TRI + 45) * total
                for l in have:
                    wp[l]["cur"]["pend"] = 0
            else:
                stats["tri"][0] += W_TRI
                stats["tri"][1] += W_TRI * len(have)
                for l in have:
                    wp[l]["cur"]["pend"] -= 1
        for l in trav0:
            c = wp[l]["cur"]
            if c["node"] == 0 and c["pend"] == 0:
                c["st"] = HIT
                wp[l]["cur"] = None
        clock[wi] += sum(v[0] for v in stats.values()) - before
    stats["_makespan"] = [max(clock), sum(clock) / n_warps]
    return stats


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_sponza_scale")
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--height", type=int, default=96)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--warps", type=int, default=48)
    a = ap.parse_args()
    px = m1.collect(a.workload, a.width, a.height, a.spp, a.depth)
    m1.W_NODE, m1.W_TRI = W_NODE, W_TRI
    base = m1.report("round-1 policy (refill 12)", m1.simulate(px, a.warps, 12))
    for K in (1, 2, 3, 4):
        for refill in (4, 8, 12):
            for t_shade in (16, 24, 28):
                for t_idle in (2, 4, 8):
                    st = simulate(px, max(1, a.warps // K), K, refill, t_shade, t_idle)
                    t = m1.report(f"K={K} refill={refill} shade>={t_shade} idle>={t_idle}", st)
    for K in (2, 3):
        st = simulate(px, max(1, a.warps // K), K, 8, 24, 4, pooled=True)
        m1.report(f"K={K} refill=8 shade>=24 idle>=4 + pooled drain", st)
