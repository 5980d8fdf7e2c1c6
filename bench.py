#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's algorithm on host cores

Workload (BASELINE.json configs[2], the scene the north-star target is quoted on): Sponza-scale
synthetic scene (~261 k triangles, textured, sky), 1920x1080, 256 spp, max depth 10, megakernel and
wavefront. One "step" = one render_frame of that configuration.

  value      Mrays/s, whole job, scene resident in HBM, device-timed (CUDA events on the launch
             stream; per-step event pairs; L2 flushed between steps; max over ranks)
  e2e        the same metric through the public API with HOST buffers: every step uploads the
             scene from host memory (rt_scene_create), builds the BVH (rt_scene_commit), renders
             and reads the RGBA8 image back to the host
  roofline   memory-bound traversal roofline of the dominant kernel: algorithmic bytes
             = rays * B(N) + samples * 16, B(N) = ceil(log8(N/4))*80 + 4*48 + 128 (+96 wavefront)
  cpu_baseline  the oracle (CPU restatement of the reference, own SAH BVH instead of Embree)
             timed on the box's host cores on a bounded sample of the same workload

N > 1 (torchrun, one rank per GPU): STRONG scaling by default — the workload's total samples per pixel are split over
the ranks (256 / N each on C3, distinct seed salts), every rank renders the full frame, and the one exchange step runs
inside the timed region: the library's fused reduce-scatter + resolve + gather kernel over NVLink peer memory
(rt_renderer_reduce_resolve: each rank sums its 1/N slice of the pixels over all ranks' accumulation buffers and stores
the resolved RGBA8 pixels into rank 0's image), ordered by two 4-byte NCCL all-reduces. --exchange nccl uses an NCCL
all-reduce of the fp32 buffers + rt_resolve instead; --sharding tiles is config 4's image-tile mode (peer-store gather
fused into the render kernel, bit-identical to one GPU); --scaling weak keeps the per-GPU work fixed (round 1's mode,
also reported under "also_weak" by the default run). The gathered image is checked, untimed, against a one-GPU render
(tiles) or against the rank-ordered sum of the accumulation buffers (spp).
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "sycl-ray-tracer_b200"

WORKLOADS = {
    # name: (scene factory name, kwargs, width, height, spp, depth)
    "c1_cube": ("cube_scene", {}, 256, 256, 1, 8),
    "c2_cornell": ("cornell_scene", {}, 1920, 1080, 64, 10),
    "c3_sponza_scale": ("sponza_scale_scene", {}, 1920, 1080, 256, 10),
    "c4_heightfield_10m": ("big_mesh_scene", {}, 3840, 2160, 16, 10),
    # configs[4]: 4K, 4096 spp TOTAL, rendered progressively (frames of --batch-spp samples chained with
    # RT_RENDER_RESUME) and split over the ranks by spp (4096 / N each): ~40 s per step on one GPU, so it is
    # an explicit --workload, not the default
    "c5_progressive_4k": ("sponza_scale_scene", {}, 3840, 2160, 4096, 10),
    # not one of BASELINE.json's configs: the BVH-hostile scene (large architectural triangles next to dense props), for the
    # tree-quality measurements of profiles/README.md
    "stadium": ("stadium_scene", {}, 1920, 1080, 64, 10),
}
PROGRESSIVE = {"c5_progressive_4k": 256}  # default --batch-spp


_JSON_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def algorithmic_bytes_per_ray(n_tris, wavefront):
    """one root-to-leaf descent of a balanced 8-wide tree with 4-triangle leaves + one
    hit's shading gather (+ the reference's queue record for the wavefront)."""
    levels = max(1, math.ceil(math.log(max(n_tris / 4.0, 1.0001), 8)))
    return levels * 80 + 4 * 48 + 128 + (96 if wavefront else 0)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region, sampled every 50 ms through NVML
    (nvidia_ml_py) on a background thread; falls back to polling nvidia-smi."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.sm, self.mx, self.reasons, self.stop_flag, self.thread, self.proc, self.path = [], [], set(), False, None, None, None
        self.slowest_query_ms = 0.0

    def _nvml_loop(self, nv, h):
        reasons = nv.nvmlDeviceGetCurrentClocksEventReasons if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                t0 = time.perf_counter()
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                t1 = time.perf_counter()
                mask = reasons(h)
                t2 = time.perf_counter()
                self.slowest_query_ms = max(self.slowest_query_ms, (t1 - t0) * 1e3, (t2 - t1) * 1e3)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode() if not str(self.uuid).startswith("GPU-") else str(self.uuid).encode())
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))  # constant: asked once, outside the loop
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread:
            if not self.sm:
                time.sleep(0.06)  # a very short timed region: make sure one sample exists
            self.stop_flag = True
            self.thread.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            try:
                for line in open(self.path):
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 6:
                        continue
                    try:
                        self.sm.append(float(c[0]))
                        self.mx.append(float(c[1]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
                os.unlink(self.path)
            except Exception:
                pass
        if self.sm:
            s = sorted(self.sm)
            out.update(sm_mhz=s[len(s) // 2], sm_max_mhz=max(self.mx), reasons=sorted(self.reasons), samples=len(s))
        return out


def build_scene_data(workload):
    scenes = importlib.import_module(PKG + ".scenes")
    fac, kw, w, h, spp, depth = WORKLOADS[workload]
    return getattr(scenes, fac)(**kw), w, h, spp, depth


def cpu_reference_run(data, w, h, depth, mode, target_s, threads=0):
    """Time the oracle (CPU restatement of the reference's trace_ray/material code, own SAH BVH in place of Embree's
    rtcIntersect1) on a bounded sample of the workload: a 6 x 4 lattice of tiles SPREAD OVER THE WHOLE FRAME (so the
    sample sees the frame's mix of sky, floor, walls and spheres, not its centre) at reduced spp (Mrays/s is
    spp-independent, benchmark_avg.csv:12-19)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    osc = _oracle.Scene(data)
    ocam = _oracle.camera_for(data, w, h)
    gx, gy = (6, 4) if w >= 480 and h >= 272 else (1, 1)
    tw, th = min(w, 80), min(h, 68)
    crops = []
    for j in range(gy):
        for i in range(gx):
            x0 = (w - tw) * i // max(1, gx - 1) if gx > 1 else (w - tw) // 2
            y0 = (h - th) * j // max(1, gy - 1) if gy > 1 else (h - th) // 2
            crops.append((x0, y0, x0 + tw, y0 + th))

    def once(spp):
        rays = secs = 0
        for c in crops:
            r = osc.render(ocam, mode, depth, spp, use_bvh=True, crop=c, threads=threads)
            rays, secs = rays + r["ray_count"], secs + r["seconds"]
        return rays, secs
    once(1)  # also builds the BVH
    rays1, secs1 = once(1)
    spp = int(max(1, min(64, round(target_s * (rays1 / max(secs1, 1e-9)) / max(rays1, 1)))))
    cores = _oracle.lib().orc_max_threads() if threads <= 0 else threads
    sample = (f"{gx}x{gy} lattice of {tw}x{th} tiles spread over the whole {w}x{h} frame ({len(crops) * tw * th} pixels), {spp} spp, depth {depth}, "
              f"{'wavefront' if mode else 'megakernel'} seeding")

    def step():
        rays, secs = once(spp)
        return rays, secs, len(crops) * tw * th * spp
    return step, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    data, w, h, spp, depth = build_scene_data(args.workload)
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    step, cores, sample = cpu_reference_run(data, w, h, depth, 0, args.cpu_seconds, threads=ncores)
    for _ in range(args.warmup):
        step()
    rays = secs = samples = 0
    for _ in range(args.steps):
        r, s, n = step()
        rays, secs, samples = rays + r, secs + s, samples + n
    val = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "msamples_per_s": samples / secs / 1e6,
        "config": {"workload": args.workload, "triangles": data.triangle_count, "width": w, "height": h,
                   "spp": spp, "max_depth": depth, "renderer": "megakernel",
                   "note": "reference algorithm on host cores: oracle C++ restatement + own SAH BVH (Embree/SYCL absent, reference unbuildable here); "
                           "each step renders the sample named in cpu_baseline.sample: tiles spread over the whole frame of this workload"},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def main():
    # libraries (NCCL's version banner) write to stdout: keep fd 1 for the ONE JSON line only
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_sponza_scale", choices=sorted(WORKLOADS))
    ap.add_argument("--renderer", default="both", choices=["both", "megakernel", "wavefront"])
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel")
    ap.add_argument("--batch-spp", type=int, default=-1, help="progressive rendering: samples per frame, frames chained with "
                    "RT_RENDER_RESUME (bit-identical to one frame); 0 = one frame, default: 256 for c5, else 0")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="target seconds per CPU-baseline step")
    ap.add_argument("--sharding", default="auto", choices=["auto", "spp", "tiles"],
                    help="N>1: spp = every rank renders the full frame with its share of the samples and its own seed salt; tiles = 64x64 "
                         "image tiles dealt round robin (config 4's mode, bit-identical to one GPU); auto = tiles for c4, spp otherwise")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1 with spp sharding: strong = the workload's total spp is split over the ranks (default); weak = every rank "
                         "renders the workload's full spp (N times the work)")
    ap.add_argument("--exchange", "--gather", dest="exchange", default="peer", choices=["peer", "nccl", "allreduce"],
                    help="N>1: peer = the library's own kernels move the result over NVLink peer memory (tiles: the render kernel stores finished "
                         "RGBA8 pixels into rank 0's image; spp: fused reduce-scatter + resolve + gather kernel), 4-byte NCCL all-reduces order "
                         "the frame; nccl = NCCL all-reduce of the fp32 accumulation buffers + rt_resolve on rank 0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the quick secondary measurements (configs[1] Cornell at N=1, the weak-scaling line at N>1)")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the untimed check of the gathered image")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.exchange == "allreduce":
        args.exchange = "nccl"
    if args.impl == "reference":
        return run_reference(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        print("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module(PKG)
    data, w, h, spp, depth = build_scene_data(args.workload)
    if args.spp:
        spp = args.spp
    app = pkg.App(local_rank)
    stream = torch.cuda.Stream()  # a non-default stream: kernels, NCCL and the timing events share it
    torch.cuda.set_stream(stream)
    app.set_stream(stream.cuda_stream)
    scene = pkg.Scene(app, data)
    stats = scene.stats
    cam = pkg.Camera((w, h), data.camera_position, data.camera_direction, data.camera_focal_length)
    sharding = args.sharding if args.sharding != "auto" else ("tiles" if args.workload == "c4_heightfield_10m" else "spp")
    tiles = sharding == "tiles" and world > 1
    batch = PROGRESSIVE.get(args.workload, 0) if args.batch_spp < 0 else args.batch_spp
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    class DevAccum:  # zero-copy view of the renderer's device accumulation buffer
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3}

    class DevImage:  # zero-copy view of a renderer's device RGBA8 image
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "|u1", "data": (int(ptr), False), "version": 3}

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    token = torch.zeros(1, dtype=torch.int32, device="cuda") if dist else None

    class Plan:
        """how one step is sharded over the ranks: samples per rank, the shard descriptor, the exchange"""

        def __init__(self, strong):
            self.strong = strong or tiles
            if world == 1:
                self.spp_rank, self.spp_total, self.shard = spp, spp, None
            elif tiles:
                self.spp_rank, self.spp_total = spp, spp
                self.shard = {"rank": rank, "world": world, "tile_size": 64, "seed_salt": 0}
            else:
                self.spp_rank = (spp // world + (1 if rank < spp % world else 0)) if strong else spp
                self.spp_total = spp if strong else spp * world
                self.shard = {"rank": rank, "world": world, "tile_size": 0, "seed_salt": (rank * 0x9E3779B9) & 0xFFFFFFFF}
            self.peer = dist is not None and args.exchange == "peer"

    def render(r, sc, n_spp, **kw):
        """one step's rendering: a single rt_render_frame, or a progressive chain of them"""
        if not batch or batch >= n_spp:
            r.sample_count = n_spp
            return r.render_frame(cam, sc, **kw)
        done, agg = 0, None
        while done < n_spp:
            r.sample_count = min(batch, n_spp - done)
            f = r.render_frame(cam, sc, resume=done > 0, **kw)
            done += r.sample_count
            if agg is None:
                agg = f
            else:
                agg.ray_count, agg.device_ms, agg.kernel_launches = agg.ray_count + f.ray_count, agg.device_ms + f.device_ms, agg.kernel_launches + f.kernel_launches
        r.sample_count = n_spp
        return agg

    def attach(r, plan):
        """peer exchange: rank 0 exports its image (gather destination of every rank), spp slices also share every rank's
        accumulation buffer (CUDA IPC handles, exchanged once per renderer)"""
        if not plan.peer:
            return
        box = [r.export_image() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            r.set_gather(handle=box[0])
        if not tiles:
            handles = [None] * world
            dist.all_gather_object(handles, r.export_accum())
            r.set_peers(handles, rank)
        barrier()

    def detach(r, plan):
        if not plan.peer:
            return
        barrier()
        if not tiles:
            r.set_peers(None)
        if rank != 0:
            r.set_gather()
        barrier()

    def exchange(r, plan, accum_t, out_rgba):
        """the one exchange step of a sharded frame, enqueued on the stream the render ran on; returns the kernels it launched here"""
        if not dist:
            return 0
        if plan.peer and tiles:      # the pixels are already in rank 0's image: order the frame across ranks
            dist.all_reduce(token)
            return 0
        if plan.peer:                # every rank has rendered -> fused reduce-scatter + resolve + gather -> rank 0's image is complete
            dist.all_reduce(token)
            r.reduce_resolve()
            dist.all_reduce(token)
            return 1
        dist.all_reduce(accum_t)     # library collective on the fp32 buffers, then resolve on rank 0
        if rank == 0:
            pkg.resolve(app, accum_t, plan.spp_total, w, h, out_rgba)
            return 1
        return 0

    def verify(r, plan, cls, accum_t, out_rgba):
        """untimed: the image the exchange produced on rank 0 against an independent computation"""
        if not dist or args.no_verify:
            return None
        barrier()
        got = torch.as_tensor(DevImage(r.device_rgba8_ptr), device="cuda").clone() if plan.peer else out_rgba.clone()
        if tiles:   # bit-identical to ONE GPU rendering the whole frame
            ok = None
            if rank == 0:
                solo = cls(app, (w, h), None, depth, spp)
                f1 = render(solo, scene, spp, want=())
                ok = bool(torch.equal(torch.as_tensor(DevImage(solo.device_rgba8_ptr), device="cuda"), got))
                solo.close()
            barrier()
            return {"image_equals_one_gpu_render": ok}
        # spp slices: the rank-ordered fp32 sum of every rank's accumulation buffer, resolved
        mine = torch.as_tensor(DevAccum(r.device_accum_ptr), device="cuda").clone() if plan.peer else None
        if mine is None:
            return None     # the NCCL all-reduce overwrote the buffers in place (and its summation order is not defined)
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        ok = None
        if rank == 0:
            total = parts[0].clone()
            for k in range(1, world):
                total += parts[k]
            want = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
            pkg.resolve(app, total, plan.spp_total, w, h, want)
            ok = bool(torch.equal(want, got)) and bool((total[..., 3] == plan.spp_total).all().item())
        barrier()
        return {"image_equals_rank_ordered_sum": ok}

    def bench_renderer(cls, name, plan, steps, warmup, check=True):
        r = cls(app, (w, h), None, depth, plan.spp_rank)
        accum_t = torch.as_tensor(DevAccum(r.device_accum_ptr), device="cuda") if dist and not plan.peer else None
        out_rgba = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
        attach(r, plan)

        def step():
            f = render(r, scene, plan.spp_rank, want=(), shard=plan.shard)
            f.kernel_launches += exchange(r, plan, accum_t, out_rgba)
            return f
        for _ in range(warmup):
            step()
        ms, kms, rays, launches = 0.0, 0.0, 0, 0
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        sampler = ClockSampler(local_rank, uuid)
        barrier()
        if rank == 0:
            sampler.start()
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            f = step()
            e1.record(stream)
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
            kms += f.device_ms
            rays += f.ray_count
            launches += f.kernel_launches
        clocks = sampler.stop() if rank == 0 else None
        barrier()
        t = torch.tensor([ms, kms, float(rays), float(launches)], dtype=torch.float64, device="cuda")
        if dist:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, kms, rays, launches = float(mx[0]), float(mx[1]), int(sm[2]), int(sm[3])
        checked = verify(r, plan, cls, accum_t, out_rgba) if check else None
        detach(r, plan)
        r.close()
        return {"name": name, "verified": checked, "ms": ms, "kernel_ms": kms, "rays": rays, "launches": launches, "clocks": clocks,
                "last_launches": f.kernel_launches, "steps": steps}

    split_total = args.workload in PROGRESSIVE and world > 1 and not tiles   # c5 is defined by its total spp
    plan = Plan(strong=(args.scaling == "strong" or split_total))
    spp_total = plan.spp_total
    results = []
    if args.renderer in ("both", "megakernel"):
        results.append(bench_renderer(pkg.MegakernelRenderer, "megakernel", plan, args.steps, args.warmup))
    if args.renderer in ("both", "wavefront"):
        results.append(bench_renderer(pkg.WavefrontRenderer, "wavefront", plan, args.steps, args.warmup))
    best = max(results, key=lambda x: x["rays"] / x["ms"])
    samples_total = w * h * spp_total * args.steps

    # ---- N > 1: the weak-scaling line next to the strong one (every rank renders the workload's full spp), quick ----
    also_weak = None
    if world > 1 and plan.strong and not tiles and not split_total and not args.no_also:
        wplan = Plan(strong=False)
        cls = pkg.MegakernelRenderer if best["name"] == "megakernel" else pkg.WavefrontRenderer
        wr = bench_renderer(cls, best["name"], wplan, 3, 3, check=False)
        also_weak = {"scaling": "weak", "spp_per_gpu": wplan.spp_rank, "renderer": best["name"], "steps": 3,
                     "mrays_per_s": wr["rays"] / (wr["ms"] * 1e-3) / 1e6, "ms_per_step": wr["ms"] / 3,
                     "msamples_per_s": w * h * wplan.spp_total * 3 / (wr["ms"] * 1e-3) / 1e6}

    # ---- the other single-GPU configuration of BASELINE.json (configs[1], Cornell, wavefront), quick ----
    also = None
    if args.workload == "c3_sponza_scale" and world == 1 and not args.no_also:
        data2, w2, h2, spp2, depth2 = build_scene_data("c2_cornell")
        sc2 = pkg.Scene(app, data2)
        cam2 = pkg.Camera((w2, h2), data2.camera_position, data2.camera_direction, data2.camera_focal_length)
        also = {"workload": "c2_cornell", "triangles": int(sc2.stats["triangle_count"]), "width": w2, "height": h2, "spp": spp2,
                "max_depth": depth2, "steps": 3}
        for cls, name in ((pkg.WavefrontRenderer, "wavefront"), (pkg.MegakernelRenderer, "megakernel")):
            r2 = cls(app, (w2, h2), None, depth2, spp2)
            for _ in range(3):
                r2.render_frame(cam2, sc2, want=())
            ms2, rays2 = 0.0, 0
            for _ in range(3):
                flush.fill_(1)
                f2 = r2.render_frame(cam2, sc2, want=())
                ms2, rays2 = ms2 + f2.device_ms, rays2 + f2.ray_count
            also[name] = {"mrays_per_s": rays2 / ms2 / 1e3, "msamples_per_s": 3 * w2 * h2 * spp2 / ms2 / 1e3, "ms_per_step": ms2 / 3}
            r2.close()
        sc2.close()

    # ---- end to end through the public API with host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        cls = pkg.MegakernelRenderer if best["name"] == "megakernel" else pkg.WavefrontRenderer
        r = cls(app, (w, h), None, depth, plan.spp_rank)
        accum_t = torch.as_tensor(DevAccum(r.device_accum_ptr), device="cuda") if dist and not plan.peer else None
        host_img = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
        out_rgba = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
        attach(r, plan)
        dev_img = torch.as_tensor(DevImage(r.device_rgba8_ptr), device="cuda") if plan.peer else out_rgba
        h2d = sum(i.positions.nbytes + i.normals.nbytes + i.uvs.nbytes + i.indices.nbytes + 64 + 40 for i in data.instances)
        h2d += (data.textures.nbytes if data.textures is not None else 0) + 56 + 24
        d2h = w * h * 4 + 8

        def e2e_step():
            t_a = time.perf_counter()
            sc = pkg.Scene(app, data, commit=False)  # H2D of the scene from host memory ...
            t_m = time.perf_counter()
            sc.commit()                               # ... + GPU BVH build
            t_b = time.perf_counter()
            e2e_step.create_s += t_b - t_a
            e2e_step.upload_s += t_m - t_a
            e2e_step.slowest_create_s = max(e2e_step.slowest_create_s, t_b - t_a)
            if dist:
                if plan.peer:
                    barrier()  # rank 0 has read the previous frame before anybody stores into its image again
                f = render(r, sc, plan.spp_rank, want=(), shard=plan.shard)
                exchange(r, plan, accum_t, out_rgba)
                if rank == 0:
                    host_img.copy_(dev_img, non_blocking=True)  # D2H of the gathered image
                    torch.cuda.current_stream().synchronize()
            else:
                f = render(r, sc, plan.spp_rank, want=("rgba8",), outputs={"rgba8": host_img})  # D2H inside (every progressive frame)
            e2e_step.render_s += time.perf_counter() - t_b
            sc.close()
            return f
        e2e_step.create_s = e2e_step.render_s = e2e_step.upload_s = e2e_step.slowest_create_s = 0.0
        e2e_step()
        e2e_step.create_s = e2e_step.render_s = e2e_step.upload_s = e2e_step.slowest_create_s = 0.0
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        e_sampler = ClockSampler(local_rank, uuid)
        barrier()
        if rank == 0:
            e_sampler.start()
        t0 = time.perf_counter()
        e_rays, e_dev_ms = 0, 0.0
        for _ in range(args.steps):
            ef = e2e_step()
            e_rays += ef.ray_count
            e_dev_ms += ef.device_ms
        barrier()
        dt = time.perf_counter() - t0
        e_clocks = e_sampler.stop() if rank == 0 else None
        t = torch.tensor([dt, float(e_rays)], dtype=torch.float64, device="cuda")
        if dist:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            dt, e_rays = float(mx[0]), int(sm[1])
        e2e = {"value": e_rays / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt / args.steps * 1e3, "render_device_ms_per_step": e_dev_ms / args.steps,
               "scene_upload_and_build_ms_per_step": e2e_step.create_s / args.steps * 1e3,
               "scene_upload_ms_per_step": e2e_step.upload_s / args.steps * 1e3, "slowest_scene_upload_and_build_ms": e2e_step.slowest_create_s * 1e3,
               "render_call_ms_per_step": e2e_step.render_s / args.steps * 1e3, "clocks": e_clocks,
               "includes": "scene upload from host (pageable numpy arrays, staged through pinned chunks) + BVH build + render"
                           + ((" + peer-memory exchange" if plan.peer else " + NCCL all-reduce + resolve") if dist else "") + " + image read-back, every step"}
        detach(r, plan)
        r.close()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = measured_peak()
    mega = next((x for x in results if x["name"] == best["name"]), best)
    wave = best["name"] == "wavefront"
    bpr = algorithmic_bytes_per_ray(stats["triangle_count"], wave)
    # per launch: megakernel = 1 launch per step; wavefront = the extend+shade pair averaged over the step's launches
    n_launch = args.steps * world  # one render kernel per rank and step in both formulations (megakernel / queue-driven wavefront)
    alg_bytes_total = mega["rays"] * bpr + samples_total * 16
    # per GPU: bytes of one rank's steps / that rank's render-kernel device time (max over ranks)
    kernel_s = mega["kernel_ms"] * 1e-3
    achieved = alg_bytes_total / world / kernel_s / 1e9
    # DRAM traffic comes from ncu (one --set full capture per workload and kernel, summarised under profiles/): recorded there as
    # bytes per ray of that capture and scaled to this run's rays per launch — a cross-reference, not something this run measured
    traffic, traffic_source = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            ent = json.load(open(tp)).get(args.workload, {}).get(best["name"], {})
            if ent.get("dram_bytes_per_ray") is not None:
                traffic = ent["dram_bytes_per_ray"] * mega["rays"] / max(1, n_launch)
                traffic_source = f"ncu capture {ent.get('capture')}: {ent['dram_bytes_per_ray']:.2f} DRAM bytes per ray x this run's rays per launch"
        except Exception:
            traffic = None
    # the realistic byte model next to the floor — visits*80 + tests*48 + 128 (+96) with the node visits and
    # triangle tests per ray counted inside the real kernels (RT_GPU_COUNTERS build, recorded in profiles/)
    realistic = None
    try:
        steps = json.load(open(tp)).get(args.workload, {}).get(best["name"], {}).get("traversal_steps")
        if steps:
            rb = steps["node_visits_per_ray"] * 80 + steps["triangle_tests_per_ray"] * 48 + 128 + (96 if wave else 0)
            realistic = {"node_visits_per_ray": steps["node_visits_per_ray"], "triangle_tests_per_ray": steps["triangle_tests_per_ray"],
                         "bytes_per_ray": rb, "achieved_gbs": (mega["rays"] * rb + samples_total * 16) / world / kernel_s / 1e9}
    except Exception:
        realistic = None
    roofline = {"bound": "hbm", "kernel": "k_megakernel" if not wave else "k_wf_flow",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                "peak_source": peak_src, "algorithmic_bytes_per_ray": bpr, "realistic_model": realistic,
                "algorithmic_bytes_per_launch": alg_bytes_total / max(1, n_launch),
                "launch_ms": mega["kernel_ms"] / max(1, n_launch / world),
                "note": ("scene (BVH+shading = %.1f MB) %s" % ((stats["bvh_bytes"] + stats["shading_bytes"]) / 1e6,
                         "is L2-resident (126 MB L2): HBM is the stated, conservative roofline; the physical limits are issue slots and the L1TEX wavefront rate (profiles/)"
                         if stats["bvh_bytes"] + stats["shading_bytes"] < 100e6 else "exceeds the 126 MB L2: node/triangle fetches are HBM traffic"))}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        step, cores, sample = cpu_reference_run(data, w, h, depth, 1 if wave else 0, args.cpu_seconds, threads=ncores)
        rr, ss, _ = step()
        rr2, ss2, _ = step()
        cpu = {"value": (rr + rr2) / (ss + ss2) / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}

    value = best["rays"] / (best["ms"] * 1e-3) / 1e6
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": best["ms"] / args.steps, "higher_is_better": True, "scaling": "strong" if plan.strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "msamples_per_s": samples_total / (best["ms"] * 1e-3) / 1e6,
        "config": {"workload": args.workload, "triangles": int(stats["triangle_count"]), "width": w, "height": h, "spp": spp,
                   "max_depth": depth, "renderer": best["name"],
                   "progressive": (f"{-(-spp // batch)} frames of {batch} spp chained with RT_RENDER_RESUME" if batch and batch < spp else None), "l2": "flushed between timed steps (256 MB write)",
                   "spp_per_gpu": plan.spp_rank if world > 1 else None,
                   "sharding": "none" if world == 1 else (
                       ("image tiles: 64x64 tiles round robin over ranks, bit-identical to 1 GPU; " if tiles else
                        f"spp slices: the frame's {spp_total} spp split over {world} ranks ({'strong' if plan.strong else 'weak'} scaling), distinct seed salts; ")
                       + (("finished RGBA8 pixels stored by the render kernel straight into rank 0's image over NVLink peer memory (CUDA IPC), 4-byte all-reduce as the frame barrier"
                           if tiles else "fused reduce-scatter + resolve + gather kernel over NVLink peer memory (rt_renderer_reduce_resolve), two 4-byte NCCL all-reduces as barriers")
                          if plan.peer else "NCCL all-reduce (sum) of the fp32 accumulation buffer + resolve on rank 0")),
                   "bvh": {"nodes": int(stats["node_count"]), "depth": int(stats["wide_depth"]), "build_ms": float(stats["build_ms"])}},
        "renderers": {x["name"]: {"mrays_per_s": x["rays"] / (x["ms"] * 1e-3) / 1e6, "ms_per_step": x["ms"] / args.steps,
                                  "msamples_per_s": samples_total / (x["ms"] * 1e-3) / 1e6,
                                  "kernel_launches_per_step": x["last_launches"], "clocks": x["clocks"],
                                  **({"verified": x["verified"]} if x.get("verified") is not None else {})} for x in results},
        "also": also, "also_weak": also_weak, "clocks": best["clocks"], "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(best["launches"]),
    }
    emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
