#!/usr/bin/env python3
"""bench.py — headline benchmark of the ray-trace hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own shader on the host cores

metric   : Mrays/s — one ray = one calcInter or inShadow evaluation; rays per frame are counted
           exactly by the instrumented kernel variant (untimed) and the count is checked against the oracle in tests.
workload : mixed1024 scene (512 spheres + 256 boxes + 192 quadrics + 64 tori), 3840x2160, 8 bounces.
build    : "fused" (FMA contraction, MUFU reciprocals, rotation matrices; parity by the envelope criterion of
           tests/envelope.py) is the headline; the "strict" build (bit-comparable with the oracle) is timed in the same run
           and reported under "strict".  --build strict swaps the roles.
step     : one full frame.  N > 1: the SAME frame, 4-scanline blocks interleaved over the ranks, one NCCL gather to
           rank 0 inside the step (strong scaling) — issued by librtb200.so itself (rtb_comm_init / rtb_gather);
           torch.distributed only ships the 128-byte NCCL id and reduces the timings.
value    : frame already resident (scene uploaded once), CUDA-event time of K steps, max over ranks.
e2e      : through the public GLWrapper API with HOST buffers: every step uploads all uniform buffers from pinned
           host memory, renders, (gathers,) and reads the RGBA32F frame back to pinned host memory — as a double-buffered
           frame loop (the read-back of frame i overlaps frame i+1; all K frames are in host memory when the clock stops).
configs  : the other BASELINE.json configs, a few frames each, in the same JSON line (N = 1: default256, default1080,
           spheres4k, tori1080; N > 1: mixed1024_8k, the 7680x4320 frame of configs[4]).
Between steps L2 is flushed by writing a 256 MiB buffer (untimed); the 133 MB frame alone exceeds the 126 MB L2.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "mixed1024_4k"
METRIC = "Mrays/s"
BLOCK_ROWS = 4            # scanlines per partition block (= the kernel's tile height): finest interleave, best balance across ranks
STRICT_OF = {"fused": 0, "strict": 1}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--build", default="fused", choices=["fused", "strict", "fast"],
                    help="fused = FMA-contracted build (headline; envelope parity); strict = no contraction, matches the oracle to 5e-7 on every pixel")
    ap.add_argument("--gather", default="cabi", choices=["cabi", "torch"],
                    help="N > 1: rtb_gather inside librtb200.so (default) or torch.distributed.gather + index_copy_ (cross-check)")
    ap.add_argument("--scale", type=float, default=1.0, help="canvas scale (development only; the contract run uses 1.0)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the other build, the culled side measurements and the other configs")
    a = ap.parse_args()
    if a.build == "fast":
        a.build = "fused"
    return a


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def sample_quads(w, h, n, seed=123):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, w // 2, n) * 2).astype(np.int32), (rng.integers(0, h // 2, n) * 2).astype(np.int32)


def bench_textures(scene):
    """The samplers both arms bind: the reference's asset files do not travel, so a procedural 512^2 cubemap (and procedural 2-D
    textures where the scene references them)."""
    from rtb200.textures import TextureSet, procedural_textures
    return procedural_textures(cube_size=512) if scene.uses_textures() else TextureSet(cube=procedural_textures(cube_size=512).cube)


def cpu_rate(scene, target_seconds, steps=1, warmup=0):
    """Times the reference's path on the host cores over a bounded, seeded sample of the workload's 2x2 quads, with the same
    samplers the GPU arm binds.  Uses the -O3 timing copy of oracle/_ref (the reference's own rt.frag compiled as C++) when it
    travelled, else the restatement built -O3 -march=native on this box; the compiler flags go into the report."""
    from oracle.binding import Oracle, Stats, have_ref, timing_flags
    w, h = int(scene.scene["canvas_width"]), int(scene.scene["canvas_height"])
    cores = os.cpu_count() or 1
    ts = bench_textures(scene)
    counter = Oracle(scene, ts)                       # the restatement counts rays (same paths, bit-identical images)
    flags = timing_flags("ref") if have_ref() else None
    if flags:
        timed, kind, what = Oracle(scene, ts, impl="ref", timing_build=True), "reference", "oracle/_ref = the reference's rt.frag compiled as C++"
    else:
        flags = timing_flags("oracle")
        timed = Oracle(scene, ts, timing_build=True) if flags else counter
        kind, what = "port", "the oracle restatement"
        flags = flags or "g++ -O2 -ffp-contract=off (portable build)"
    qx, qy = sample_quads(w, h, 256 * cores, seed=7)  # calibration
    t0 = time.perf_counter()
    timed.render_quads(qx, qy, threads=cores)
    per_quad = (time.perf_counter() - t0) / len(qx)
    n = int(max(256, min(w * h // 4, target_seconds / max(per_quad, 1e-9))))
    qx, qy = sample_quads(w, h, n)
    st = Stats()
    counter.render_quads(qx, qy, threads=cores, stats=st)
    rays = st.rays_nearest + st.rays_shadow
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        timed.render_quads(qx, qy, threads=cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = float(np.mean(times))
    return {"value": rays / mean / 1e6, "unit": METRIC, "cores": cores, "kind": kind, "flags": flags,
            "sample": f"{n} seeded random 2x2 quads ({4 * n} px, {rays} rays) of the {w}x{h} frame per step, all {cores} host threads, same procedural "
                      f"cubemap as the GPU arm; timed on {what}, built with `{flags}`",
            "seconds_per_step": mean}, rays, mean


def describe(scene):
    """Workload facts for the JSON line, from the scene itself."""
    d = scene.get_defines()
    counts = {k: int(d[k + "_size"]) for k in ("sphere", "plane", "surface", "box", "torus", "ring", "light_point", "light_direct")}
    prims = sum(counts[k] for k in ("sphere", "plane", "surface", "box", "torus", "ring"))
    text = " + ".join(f"{v} {k}" for k, v in counts.items() if v and not k.startswith("light")) + f", {counts['light_point'] + counts['light_direct']} lights"
    return text, prims, int(d["iterations"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rtb200  # noqa: F401
    from rtb200 import scenes
    scene = scenes.build_config(args.workload, args.scale)
    base, rays, mean = cpu_rate(scene, args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
    w, h = int(scene.scene["canvas_width"]), int(scene.scene["canvas_height"])
    _, n_prims, n_bounces = describe(scene)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "width": w, "height": h, "primitives": n_prims, "bounces": n_bounces, "skybox": "procedural 512^2 cubemap",
                       "note": "reference = the repo's GLSL shader executed on the host CPU cores (no GL available in this image); "
                               "each step renders a bounded seeded sample of the frame's quads"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "flags", "sample")},
            "e2e": {"value": base["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
class Workload:
    """One BASELINE config set up on this rank's GPU through the public GLWrapper API, ready to be timed."""

    def __init__(self, name, scale, rank, local_rank, world, gather_mode):
        import torch
        import torch.distributed as dist
        import rtb200
        from rtb200 import dist as rdist, scenes
        self.torch, self.dist, self.rdist = torch, dist, rdist
        self.name, self.rank, self.world, self.gather_mode = name, rank, world, gather_mode
        self.dev = torch.device("cuda", local_rank)
        self.scene = scenes.build_config(name, scale)
        self.w, self.h = int(self.scene.scene["canvas_width"]), int(self.scene.scene["canvas_height"])
        w, h = self.w, self.h
        self.gl = gl = rtb200.GLWrapper(w, h, False, device=local_rank)
        gl.init_window()
        if world > 1:
            if gather_mode == "cabi":
                ids = [gl.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)            # the launcher's plumbing ships 128 bytes; the data path is rtb_gather (NCCL inside the library)
                gl.comm_init(ids[0], rank, world, BLOCK_ROWS)
            else:
                gl.set_partition(rank, world, BLOCK_ROWS)
        # host-side inputs live in pinned memory (e2e uploads them every step)
        arrays = {name_: self.scene.array(attr) for name_, attr in (("spheres_buf", "spheres"), ("planes_buf", "planes"), ("surfaces_buf", "surfaces"),
                  ("boxes_buf", "boxes"), ("toruses_buf", "toruses"), ("rings_buf", "rings"), ("lights_point_buf", "lights_point"),
                  ("lights_direct_buf", "lights_direct"))}
        arrays["scene_buf"] = np.ascontiguousarray(self.scene.scene).reshape(1)
        self.pinned = {}
        for name_, a in arrays.items():
            t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
            t[: a.nbytes] = torch.from_numpy(np.frombuffer(a.tobytes(), dtype=np.uint8).copy())
            self.pinned[name_] = t.numpy()[: a.nbytes]
        self.h2d_bytes = sum(v.nbytes for v in self.pinned.values())
        self.handles = rtb200.setup_scene(gl, self.scene, bench_textures(self.scene))
        pad_rows = rdist.max_local_rows(h, world, BLOCK_ROWS)
        self.local = torch.zeros((pad_rows, w, 4), dtype=torch.float32, device=self.dev)
        root_multi = rank == 0 and world > 1
        self.full = torch.empty((h, w, 4), dtype=torch.float32, device=self.dev) if root_multi else None
        self.scratch = [torch.empty_like(self.local) for _ in range(world)] if (root_multi and gather_mode == "torch") else None
        self.host_frame = torch.empty((h, w, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)   # a real (non-legacy) stream: kernels, NCCL and the timing events all use it
        torch.cuda.set_stream(self.stream)
        self.local_ms = 0.0

    def close(self):
        self.gl.stop()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def render_only(self):
        self.gl.draw_to(self.local.data_ptr(), self.stream.cuda_stream)

    def render_step(self, local=None, full=None):
        """one frame: every rank renders its blocks, the root ends up with the whole frame"""
        local = self.local if local is None else local
        full = self.full if full is None else full
        self.gl.draw_to(local.data_ptr(), self.stream.cuda_stream)
        if self.world == 1:
            return local
        if self.gather_mode == "cabi":
            self.gl.gather(local.data_ptr(), full.data_ptr() if self.rank == 0 else 0, self.stream.cuda_stream)
            return full
        return self.rdist.gather_frame(local, self.h, self.rank, self.world, BLOCK_ROWS, out=full, scratch=self.scratch)

    def timed_loop(self, fn, steps, warm):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps; max over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in ev:
            self.flush.fill_(1)
            a.record(self.stream)
            fn()
            b.record(self.stream)
        self.barrier()
        self.local_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        ms = torch.tensor([self.local_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def count(self):
        """exact work of this rank's share (instrumented variant, untimed), summed over ranks: rays, flops, DK trips, pixels"""
        st = self.gl.draw_counted()
        cnt = self.torch.tensor([st.rays_nearest + st.rays_shadow, st.flops, st.dk_iterations, st.pixels], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(cnt)
        return [float(x) for x in cnt.tolist()], float(st.flops)

    def e2e(self, steps, warm):
        """host buffers in, host frame out, wall clock, max over ranks.  A double-buffered frame loop: the device->host copy of
        frame i runs on a copy stream while frame i+1 is uploaded and rendered into the other buffer; a step ends when frame
        i-1 has arrived in pinned host memory, the timed region ends when the last frame has."""
        torch = self.torch
        if not hasattr(self, "slots"):
            self.slots = [(self.local, self.full, self.host_frame),
                          (torch.zeros_like(self.local), torch.empty_like(self.full) if self.full is not None else None,
                           torch.empty_like(self.host_frame).pin_memory() if self.host_frame is not None else None)]
            self.copy_stream = torch.cuda.Stream(device=self.dev)
            self.ev_frame = [torch.cuda.Event() for _ in range(2)]
            self.ev_copied = [torch.cuda.Event() for _ in range(2)]
        self.e2e_frames_delivered = 0

        def step(i):
            s = i & 1
            local, full, host = self.slots[s]
            for name_, a in self.pinned.items():
                if a.nbytes:
                    self.gl.update_buffer(self.handles[name_], a)
            self.stream.wait_event(self.ev_copied[s])              # frame i-2 has left this slot's device buffers
            out = self.render_step(local, full)
            if self.rank == 0:
                self.ev_frame[s].record(self.stream)
                self.copy_stream.wait_event(self.ev_frame[s])
                with torch.cuda.stream(self.copy_stream):
                    host.copy_(out if self.world > 1 else out[: self.h], non_blocking=True)
                    self.ev_copied[s].record(self.copy_stream)
                if i > 0:
                    self.ev_copied[1 - s].synchronize()             # frame i-1 is in host memory
                    self.e2e_frames_delivered += 1

        def run(n):
            for i in range(n):
                step(i)
            if self.rank == 0:
                self.ev_copied[(n - 1) & 1].synchronize()
                self.e2e_frames_delivered += 1
            torch.cuda.synchronize()

        run(max(2, warm))
        self.barrier()
        self.e2e_frames_delivered = 0
        t0 = time.perf_counter()
        run(steps)
        self.barrier()
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / steps], dtype=torch.float64, device=self.dev)
        assert self.rank != 0 or self.e2e_frames_delivered == steps
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def rank_list(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            allv = [self.torch.zeros_like(t) for _ in range(self.world)]
            self.dist.all_gather(allv, t)
            t = self.torch.cat(allv)
        return [round(float(x), 3) for x in t.tolist()]


def config_line(name, rank, local_rank, world, gather_mode, peak, build, steps=4, warm=2):
    """value / ms / roofline fraction / e2e of one of the other BASELINE configs (a few frames), both builds"""
    wl = Workload(name, 1.0, rank, local_rank, world, gather_mode)
    out = {"width": wl.w, "height": wl.h, "n_gpus": world}
    try:
        for b in (build, "strict" if build == "fused" else "fused"):
            wl.gl.set_option("strict", STRICT_OF[b])
            (rays, flops, _, _), local_flops = wl.count()
            ms = wl.timed_loop(wl.render_step, steps, warm)
            ms_k = wl.timed_loop(wl.render_only, steps, 1)
            entry = {"value": rays / (ms * 1e-3) / 1e6, "ms_per_step": ms, "kernel_ms": ms_k, "rank_kernel_ms": wl.rank_list(wl.local_ms),
                     "frac": (local_flops / (ms_k * 1e-3) / 1e12 / peak) if peak else None, "kernel": int(wl.gl.stats().kernel_used)}
            if b == build:
                e2e_ms = wl.e2e(steps, 1)
                entry["e2e"] = {"value": rays / (e2e_ms * 1e-3) / 1e6, "ms_per_step": e2e_ms}
                out.update(entry)
                out["build"] = b
                out["rays_per_frame"] = rays
            else:
                out[b] = entry
    finally:
        wl.close()
    return out


def run_ours(args):
    import torch
    import rtb200
    from rtb200 import dist as rdist

    rank, local_rank, world = rdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the ray-trace pass has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist

    wl = Workload(args.workload, args.scale, rank, local_rank, world, args.gather)
    gl, w, h = wl.gl, wl.w, wl.h
    other = "strict" if args.build == "fused" else "fused"
    gl.set_option("strict", STRICT_OF[args.build])

    # ---- exact work count (instrumented variant, untimed) ----
    (rays, flops, dk_iters, pixels), local_flops = wl.count()

    # ---- value: resident inputs, device time ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = wl.timed_loop(wl.render_step, args.steps, args.warmup)
    clocks = sampler.result()
    value = rays / (ms_step * 1e-3) / 1e6

    # ---- kernel alone (roofline) on this rank ----
    ms_kernel = wl.timed_loop(wl.render_only, max(3, args.steps), 1)
    kstats = gl.stats()
    rank_kernel_ms = wl.rank_list(wl.local_ms)

    # ---- e2e: host buffers in, host frame out, wall clock ----
    e2e_ms = wl.e2e(args.steps, min(args.warmup, 2))
    e2e_value = rays / (e2e_ms * 1e-3) / 1e6
    checksum = float(wl.host_frame[..., :3].double().mean()) if rank == 0 else 0.0

    peak = rtb200.measure_fp32_peak(local_rank)
    peak3 = rtb200.measure_fp32_peak(local_rank, three_registers=True)
    extras = {}
    if not args.no_extras:
        # the other build, same run: value, kernel time, roofline fraction, per-rank kernel times
        gl.set_option("strict", STRICT_OF[other])
        (_, _, _, _), o_flops = wl.count()
        o_ms = wl.timed_loop(wl.render_step, max(2, args.steps // 2), 1)
        o_msk = wl.timed_loop(wl.render_only, max(2, args.steps // 2), 1)
        extras[other] = {"value": rays / (o_ms * 1e-3) / 1e6, "ms_per_step": o_ms, "kernel_ms": o_msk, "rank_kernel_ms": wl.rank_list(wl.local_ms),
                         "roofline_frac": o_flops / (o_msk * 1e-3) / 1e12 / peak if peak else None,
                         "parity": "bit-comparable with the oracle: every pixel within 1e-4 (observed 5e-7), work counters equal" if other == "strict"
                                   else "envelope criterion (tests/envelope.py)"}
        for label, opts in ((args.build + "_culled", {"strict": STRICT_OF[args.build], "cull": 1}),):
            for k, v in opts.items():
                gl.set_option(k, v)
            ms = wl.timed_loop(wl.render_step, max(2, args.steps // 2), 1)
            extras[label] = {"value": rays / (ms * 1e-3) / 1e6, "ms_per_step": ms,
                             "note": "conservative bounding-sphere reject before the torus solve; same image; never enters roofline.achieved"}
        gl.set_option("strict", STRICT_OF[args.build])
        gl.set_option("cull", 0)
    # ---- the SMAA post-pass on this frame: four kernels behind the ray-trace pass ----
    smaa = None
    if rank == 0 and world == 1 and not args.no_extras:
        from rtb200.textures import smaa_tables
        tabs = smaa_tables()
        if tabs is not None:
            gl.smaa_set_tables(*tabs)
            gl.enable_SMAA(3)                                    # ULTRA, main.cpp:32
            times = []
            for _ in range(5):
                wl.flush.fill_(1)
                gl.draw()                                        # ray trace + quantise + edge / weights / neighbourhood passes
                times.append(gl.smaa_last_ms())
            gl.enable_SMAA(None)
            ms_smaa = float(np.median(times))
            peaks_ = {}
            try:
                peaks_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm = peaks_.get("hbm_gbs", 6650.0)
            alg = 24.0 * w * h                                   # B per pixel: pass 1 reads 4 writes 2, pass 2 reads 2 writes 4, pass 3 reads 4 + 4 writes 4
            smaa = {"preset": "ULTRA", "ms": ms_smaa, "kernels": 4, "algorithmic_bytes": alg, "achieved_GBs": alg / (ms_smaa * 1e-3) / 1e9, "peak_GBs": hbm,
                    "frac": alg / (ms_smaa * 1e-3) / 1e9 / hbm, "share_of_frame": ms_smaa / (ms_kernel + ms_smaa),
                    "note": "HBM roofline of the post-pass alone (edge / classify / weights-over-the-compacted-edge-pixels / neighbourhood).  All four "
                            "are bound by instruction issue, not by HBM: the sampler the shader relies on (bilinear fetches at coordinates that are "
                            "a rounding error off the texel centres) is reproduced in fp32 so that the 8-bit outputs are bit-identical to the "
                            "reference's SMAA.h compiled as C++ (tests/test_smaa.py)"}
    h2d_bytes = wl.h2d_bytes
    wl.close()

    configs = {}
    if not args.no_extras and args.workload == WORKLOAD and args.scale == 1.0:
        names = ("default256", "default1080", "spheres4k", "tori1080") if world == 1 else ("mixed1024_8k",)
        for name in names:
            configs[name] = config_line(name, rank, local_rank, world, args.gather, peak, args.build)

    if rank == 0:
        from rtb200 import scenes
        scene = scenes.build_config(args.workload, args.scale)
        scene_text, n_prims, n_bounces = describe(scene)
        achieved = local_flops / (ms_kernel * 1e-3) / 1e12
        alg_bytes = (h * w * 16) / world + h2d_bytes
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        base, _, _ = cpu_rate(scene, args.cpu_seconds)
        traffic, ncu_note, executed = None, None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}_{args.build}", {})
            traffic = prof.get("dram_bytes_per_launch")
            ncu_note = {k: v for k, v in prof.items() if k != "dram_bytes_per_launch"}
            if prof.get("executed_fp32_flops") and world == 1:
                executed = {"flops_per_launch": prof["executed_fp32_flops"], "TFLOPs": prof["executed_fp32_flops"] / (ms_kernel * 1e-3) / 1e12,
                            "frac_of_peak": prof["executed_fp32_flops"] / (ms_kernel * 1e-3) / 1e12 / peak if peak else None,
                            "note": "fp32 operations the FMA pipe really executed (ncu SASS opcode counts x live lanes: FFMA 2, FMUL/FADD 1) "
                                    "over this run's kernel time: shows how much of `achieved` is strength reduction (rotation matrices "
                                    "instead of quaternion sandwiches) rather than pipe utilisation"}
        except Exception:
            pass
        # kernels of ours per timed step: every rank's ray-trace kernel, its three tile-order kernels (persistent kernel, >= 4096 tiles:
        # option "lpt"), and the root's de-interleave pass behind the NCCL fan-in
        lpt_on = kstats.kernel_used == 2 and ((w + 7) // 8) * ((h // world + 3) // 4) >= 4096
        n_launch = args.steps * (world * (4 if lpt_on else 1) + (1 if world > 1 and args.gather == "cabi" else 0))
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": scene_text + " (SURVEY.md 8d generator)",
                       "width": w, "height": h, "primitives": n_prims, "bounces": n_bounces, "build": args.build, "skybox": "procedural 512^2 cubemap",
                       "parity": "envelope criterion: no determined pixel beyond 1e-4 of the fp32 oracle (tests/test_envelope.py)" if args.build == "fused"
                                 else "every pixel within 1e-4 of the fp32 oracle (observed 5e-7)",
                       "parallelism": (f"rowblock{BLOCK_ROWS}x{world}+" + ("nccl-gather(librtb200)" if args.gather == "cabi" else "torch-gather")) if world > 1 else "single",
                       "l2": "flushed between steps (256 MiB write); the 133 MB frame exceeds L2", "kernel": int(kstats.kernel_used),
                       "grid": int(kstats.grid), "block": int(kstats.block), "smem_bytes": int(kstats.smem_bytes)},
            "rays_per_frame": rays, "pixels": pixels, "rank_kernel_ms": rank_kernel_ms, "dk_iterations": dk_iters, "frame_checksum": checksum,
            "e2e": {"value": e2e_value, "unit": METRIC, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d_bytes) * world,
                    "d2h_bytes_per_step": int(h * w * 16),
                    "loop": "double-buffered: the D2H copy of frame i (copy stream) overlaps the upload + kernel of frame i+1; wall clock from the first "
                            "upload until the last of the K frames is in pinned host memory"},
            "gpu_launches": int(n_launch),
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "ncu": ncu_note, "executed": executed,
                         "kernel": "persistent_kernel" if kstats.kernel_used == 2 else "quad_kernel",
                         "kernel_ms": ms_kernel, "algorithmic_flops_per_launch": local_flops,
                         "operand_limit": {"three_register_ffma_TFLOPs": peak3, "frac_of_it": achieved / peak3 if peak3 else None,
                                           "note": "the same FFMA chains as `peak`, but every FFMA reads three distinct registers (acc = p*q + acc) instead of one "
                                                   "register + a uniform register + a reused operand: the most a sub-partition's register file delivers to FMAs "
                                                   "that combine three live values, as the Durand-Kerner steps do (tools/micro/rf_probe.cu, "
                                                   "profiles/r2_rf_probe.txt).  Context only: `frac` is quoted against `peak`."},
                         "peak_source": "FFMA microbenchmark measured in this run (rtb_measure_fp32_peak); MEASURED_PEAKS.json has no fp32 entry; "
                                        "nominal 74.4 = 148 SM x 128 lanes x 2 x 1.965 GHz",
                         "note": "fp32 CUDA-core bound (no dense contraction, north_star).  `achieved` = ALGORITHMIC flops (SURVEY.md 8d constants x exact "
                                 "counters of this build's own rays) / kernel time; `executed` = what the pipe really did.  The strict build rounds every "
                                 "multiply and add separately (bit parity with the shader): its ceiling against an FFMA peak is 0.5",
                         "hbm": {"achieved_GBs": alg_bytes / (ms_kernel * 1e-3) / 1e9, "peak_GBs": hbm_peak,
                                 "frac": alg_bytes / (ms_kernel * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes}},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "flags", "sample")},
        }
        line.update(extras)
        if smaa:
            line["smaa"] = smaa
        if configs:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # (At N > 1 this image's NCCL_DEBUG=VERSION makes NCCL itself print a "NCCL version ..." banner to stdout before the JSON line; it is
    # left alone on purpose — NCCL's log level and destination belong to whoever launches the benchmark.)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
