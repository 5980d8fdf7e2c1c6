"""Development check run on the B200 box: parity vs the oracle + rough timing of both renderers."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
pkg = importlib.import_module("sycl-ray-tracer_b200")
scenes = importlib.import_module("sycl-ray-tracer_b200.scenes")
import _oracle

def cmp_hits(a, b, name):
    same = (a['inst'] == b['inst']) & (a['prim'] == b['prim'])
    bit = same & (a['t'].view(np.uint32) == b['t'].view(np.uint32)) & (a['u'].view(np.uint32) == b['u'].view(np.uint32)) & (a['v'].view(np.uint32) == b['v'].view(np.uint32))
    print(f"  {name}: n={len(same)} hits={int((a['inst']>=0).sum())} id-equal={same.mean():.6f} bit-equal={bit.mean():.6f}", flush=True)

app = pkg.App(0)
print("device:", app.device_name, flush=True)
rs = np.random.RandomState(1)
for name, d, ext, (w, h) in [("cube", scenes.cube_scene(), 4.0, (256, 256)), ("soup", scenes.random_soup(500, 3, 1.0, 3), 2.0, (64, 48)),
                             ("cornell", scenes.cornell_scene(3), 1.5, (96, 64))]:
    sc = pkg.Scene(app, d); osc = _oracle.Scene(d)
    print(name, sc.stats, flush=True)
    org = ((rs.rand(50000, 3) - 0.5) * 2 * ext).astype(np.float32); dr = (rs.rand(50000, 3) - 0.5).astype(np.float32)
    g = pkg.intersect(app, sc, org, dr); o = osc.intersect(org, dr, use_bvh=True)
    cmp_hits(o, g, "oracle-vs-gpu intersect")
    cam = pkg.Camera((w, h), d.camera_position, d.camera_direction, d.camera_focal_length)
    ocam = _oracle.camera_for(d, w, h)
    for cls, mode in ((pkg.MegakernelRenderer, 0), (pkg.WavefrontRenderer, 1)):
        r = cls(app, (w, h), None, 8, 4)
        f = r.render_frame(cam, sc); oo = osc.render(ocam, mode, 8, 4, use_bvh=True)
        print(f"  {cls.__name__}: rays {f.ray_count} vs {oo['ray_count']}  accum-biteq {(f.accum.view(np.uint32)==oo['accum'].view(np.uint32)).all(-1).mean():.6f} "
              f"rgba-eq {(f.rgba8==oo['rgba8']).all(-1).mean():.6f} rng-eq {(f.rng_state==oo['rng_state']).mean():.6f} ms {f.device_ms:.3f} launches {f.kernel_launches}", flush=True)
        r.close()
    sc.close()

# timing on the ~260k scene, 1080p
t = time.time(); d = scenes.sponza_scale_scene(); print("gen sponza_scale", d.triangle_count, "tris", time.time() - t, flush=True)
t = time.time(); sc = pkg.Scene(app, d); print("scene create+commit", time.time() - t, sc.stats, flush=True)
w, h = 1920, 1080
cam = pkg.Camera((w, h), d.camera_position, d.camera_direction, d.camera_focal_length)
for cls in (pkg.MegakernelRenderer, pkg.WavefrontRenderer):
    for spp in (4, 16):
        r = cls(app, (w, h), None, 10, spp)
        for it in range(2):
            f = r.render_frame(cam, sc, want=())
        print(f"  {cls.__name__} spp={spp}: {f.ray_count} rays {f.device_ms:.2f} ms -> {f.ray_count/f.device_ms/1e3:.1f} Mrays/s, {w*h*spp/f.device_ms/1e3:.1f} Msamples/s, launches {f.kernel_launches}", flush=True)
        r.close()
# parity on a crop of the big scene
osc = _oracle.Scene(d); ocam = _oracle.camera_for(d, w, h)
crop = (900, 500, 964, 532)
for cls, mode in ((pkg.MegakernelRenderer, 0), (pkg.WavefrontRenderer, 1)):
    r = cls(app, (w, h), None, 10, 2); f = r.render_frame(cam, sc)
    oo = osc.render(ocam, mode, 10, 2, use_bvh=True, crop=crop)
    x0, y0, x1, y1 = crop
    print(f"  {cls.__name__} crop parity: accum-biteq {(f.accum[y0:y1,x0:x1].view(np.uint32)==oo['accum'].view(np.uint32)).all(-1).mean():.6f} rng-eq {(f.rng_state[y0:y1,x0:x1]==oo['rng_state']).mean():.6f}", flush=True)
    r.close()
print("done", flush=True)
