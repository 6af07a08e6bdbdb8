#!/usr/bin/env python
"""Benchmark of the mip-pyramid hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): full mip chain of a synthetic
16384x16384 sRGBA8 image, 15 levels, 1 431 655 764 algorithmic bytes (level 0 read once + every other level
written once).  Level 0 is the Julia-set texture the reference demo mip-maps every frame at this very size
(--input julia, default); uniform random bytes (the worst case for shared-memory bank conflicts) and a smooth
gradient are timed in the same run and reported under config.other_inputs.  A "step" is one full-chain generation.  Two distinct 1.43 GB chains are alternated, each far
larger than the 126 MB L2, so level 0 is never cache resident.  At N > 1 every rank runs the same step on its own
image (independent units, no data-path collective): weak scaling, value = N * bytes / max-over-ranks time.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU generator
(cpuGenerateMipmaps_sRGBA compiled in place from /root/reference into oracle/_ref, else our C port of it) on
the host cores; that leg and the `cpu_baseline` object are the only places this file touches oracle/.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 16384
METRIC = "mip_chain_effective_GBps_16384x16384_srgba8"
UNIT = "GB/s"


def algorithmic_bytes(w, h, first_level=0, last_level=None, bpt=4):
    total, lvl = 0, 0
    while True:
        lw, lh = max(1, w >> lvl), max(1, h >> lvl)
        if lvl >= first_level and (last_level is None or lvl <= last_level):
            total += lw * lh * bpt
        if lw == 1 and lh == 1:
            break
        lvl += 1
    return total


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        return [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]

    def stop(self):
        if self.proc:
            self.proc.terminate()

    @staticmethod
    def summarise(rows):
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 4 + i and r[4 + i].lower().startswith("active")
                                                         for r in rows)]
        mx = [int(float(r[1])) for r in rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# ----------------------------------------------------------------------------- CPU arms
def _load_cpu_generator():
    """('reference', fn) from oracle/_ref when present, else ('port', fn) from our C restatement."""
    P = C.c_void_p
    ref = os.path.join(ROOT, "oracle", "_ref", "libnvpyr_ref.so")
    if os.path.exists(ref):
        lib = C.CDLL(ref)
        lib.ref_storage_create.restype = P
        lib.ref_storage_create.argtypes = [C.c_uint32, C.c_uint32, P]
        lib.ref_storage_generate.argtypes = [P]
        lib.ref_storage_destroy.argtypes = [P]

        def make(level0, w, h):
            s = lib.ref_storage_create(w, h, level0.ctypes.data)
            return (lambda: lib.ref_storage_generate(s)), (lambda: lib.ref_storage_destroy(s))
        return "reference", make
    so = os.path.join(ROOT, "oracle", "libnvpyr_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    lib = C.CDLL(so)
    lib.nvo_cpu_chain.argtypes = [C.c_int, P, C.c_uint32, C.c_uint32]
    lib.nvo_chain_texels.restype = C.c_uint64

    def make(level0, w, h):
        import numpy as np
        buf = np.zeros(4 * lib.nvo_chain_texels(w, h, lib.nvo_level_count(w, h)), dtype=np.uint8)
        buf[:level0.size] = level0
        return (lambda: lib.nvo_cpu_chain(0, buf.ctypes.data, w, h)), (lambda: None)
    return "port", make


def cpu_baseline_single(sample_edge=16384):
    """The reference's CPU generator on ONE host thread (how the reference itself runs it, one std::thread
    per image, demo_app/mipmaps_app.cpp:651-652) over a bounded sample of the workload."""
    import numpy as np
    kind, make = _load_cpu_generator()
    rng = np.random.default_rng(0)
    l0 = rng.integers(0, 256, 4 * sample_edge * sample_edge, dtype=np.uint8)
    run, free = make(l0, sample_edge, sample_edge)
    t = time.perf_counter()
    run()
    dt = time.perf_counter() - t
    free()
    by = algorithmic_bytes(sample_edge, sample_edge)
    return {"value": by / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"one {sample_edge}x{sample_edge} sRGBA8 full chain ("
                      + ("the whole workload of one step" if sample_edge == W else
                         f"1/{(W // sample_edge) ** 2} of the 16384^2 workload's texels")
                      + f"), {dt:.2f} s on one host thread"}


def run_reference_arm(args):
    """--impl reference: the reference CPU generator, one image per host thread, all host cores."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, make = _load_cpu_generator()
    try:
        cores = len(os.sched_getaffinity(0))  # the host threads this process may actually use
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 256))
    edge = 2048  # per-thread sample image; cores * steps of them stay within a few minutes
    rng = np.random.default_rng(0)
    jobs = [make(rng.integers(0, 256, 4 * edge * edge, dtype=np.uint8), edge, edge) for _ in range(cores)]

    def step():
        ts = [threading.Thread(target=j[0]) for j in jobs]
        [t.start() for t in ts]
        [t.join() for t in ts]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    [j[1]() for j in jobs]
    by = algorithmic_bytes(edge, edge) * cores * args.steps
    value = by / dt / 1e9
    sample = (f"each step = {cores} independent {edge}x{edge} sRGBA8 full chains, one per host thread "
              f"(the reference's own thread-per-image parallelism); same bytes-per-texel metric as the 16384^2 workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 16384x16384 sRGBA8 full mip chain (BASELINE configs[2]); CPU arm runs a "
                               "bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- synthetic inputs
def fill_julia(level0, w, h):
    """Julia-set texture of the reference demo (shaders/julia.comp:27-63 with demo_app/julia.cpp:65-81's
    constants, maxIterations 64) written into level0 (uint8 [h, w, 4] CUDA view), tile by tile."""
    import math
    import torch
    alpha = 2109710467 * 1.4629180792671596e-09
    cr, ci = float(0.7885 * math.sin(alpha)), float(0.7885 * math.cos(alpha))
    scale, off_r, off_i, max_it = 4.0 / w, -2.0, 2.0 * h / w, 64
    dev = level0.device
    rows = 1024
    xs = torch.arange(w, device=dev, dtype=torch.float32) * scale + off_r
    for y0 in range(0, h, rows):
        ys = torch.arange(y0, min(h, y0 + rows), device=dev, dtype=torch.float32) * -scale + off_i
        zr = xs[None, :].expand(ys.numel(), w).clone()
        zi = ys[:, None].expand(ys.numel(), w).clone()
        it = torch.zeros_like(zr, dtype=torch.int32)
        alive = torch.ones_like(zr, dtype=torch.bool)
        for _ in range(max_it):
            alive &= (zr * zr + zi * zi) <= 4
            tr = zr * zr - zi * zi + cr
            ti = 2 * zr * zi + ci
            zr = torch.where(alive, tr, zr)
            zi = torch.where(alive, ti, zi)
            it += alive.to(torch.int32)
        itf = it.to(torch.float32)
        low = it < 16
        s = (4 + itf) * 0.05
        n = torch.clamp(127.0 * (itf - 16) / torch.clamp(max_it - itf, min=1.0), 0, 255).floor()
        out = level0[y0:y0 + ys.numel()]
        out[..., 0] = torch.where(low, torch.zeros_like(s), n).to(torch.uint8)
        out[..., 1] = torch.where(low, 128 * s, 128 + torch.floor(n / 4)).to(torch.uint8)
        out[..., 2] = torch.where(low, 255 * s, 255 - n).to(torch.uint8)
        out[..., 3] = torch.where(low, 255 * s, torch.full_like(s, 255)).to(torch.uint8)


def fill_gradient(level0, w, h):
    """Opaque smooth gradient with a little dither (low-entropy case)."""
    import torch
    dev = level0.device
    x = torch.arange(w, device=dev, dtype=torch.float32)[None, :]
    for y0 in range(0, h, 2048):
        y = torch.arange(y0, min(h, y0 + 2048), device=dev, dtype=torch.float32)[:, None]
        out = level0[y0:y0 + y.numel()]
        d = ((x.to(torch.int32) * 7 + y.to(torch.int32) * 13) & 3).to(torch.float32)
        out[..., 0] = (x * (255.0 / w) + d * 0.25).expand(y.numel(), w).to(torch.uint8)
        out[..., 1] = (y * (255.0 / h) + d * 0.25).expand(y.numel(), w).to(torch.uint8)
        out[..., 2] = (127.5 + 127.5 * torch.sin(x / 97.0) * torch.cos(y / 131.0)).to(torch.uint8)
        out[..., 3] = 255


# ----------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's affinity mask) BEFORE any pinned host
    memory is allocated, so that the staging buffers of the end-to-end leg live on the GPU's own NUMA node: with
    one process per GPU the host side of the round trip otherwise crosses the socket interconnect for half of the
    ranks.  Best effort: returns the number of CPUs bound to, or 0."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [64 * i + b for i, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import vk_compute_mipmaps_b200 as nv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pipes = nv.PyramidPipelines()
    chain_bytes = nv.chain_bytes(W, H)
    assert chain_bytes == algorithmic_bytes(W, H) == 1431655764
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    bufs = []
    for _ in range(2):  # two distinct chains, alternated (each >> L2)
        b = torch.empty(chain_bytes, dtype=torch.uint8, device=dev)
        b[:4 * W * H] = torch.randint(0, 256, (4 * W * H,), dtype=torch.uint8, device=dev, generator=gen)
        bufs.append(b)
    def fill_input(name):
        if name == "random":
            for b in bufs:
                b[:4 * W * H] = torch.randint(0, 256, (4 * W * H,), dtype=torch.uint8, device=dev, generator=gen)
        else:
            {"julia": fill_julia, "gradient": fill_gradient}[name](bufs[0][:4 * W * H].view(H, W, 4), W, H)
            bufs[1][:4 * W * H].copy_(bufs[0][:4 * W * H])
    if args.input != "random":
        fill_input(args.input)
    stream = torch.cuda.current_stream()

    def step(i):
        nv.cmd_pyramid_dispatch(stream, pipes, W, H, image=bufs[i & 1])

    # ---- whole-chain timing (value) ----
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    launches0 = nv.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = nv.launch_count() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = world * chain_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- dominant kernel alone: the 6-level fast kernel on level 0 (levelCount = 7 -> exactly one launch) ----
    k_bytes = algorithmic_bytes(W, H, 0, 6)
    for i in range(3):
        nv.cmd_pyramid_dispatch(stream, pipes, W, H, 7, image=bufs[i & 1])
    torch.cuda.synchronize()
    l0 = nv.launch_count()
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev0.record(stream)
    for i in range(args.steps):
        nv.cmd_pyramid_dispatch(stream, pipes, W, H, 7, image=bufs[i & 1])
    kev1.record(stream)
    torch.cuda.synchronize()
    assert nv.launch_count() - l0 == args.steps
    k_ms = kev0.elapsed_time(kev1) / args.steps

    # ---- same chain on the other synthetic inputs of SURVEY 8d (bytes moved are identical) ----
    other_inputs = {}
    if rank == 0 and not args.no_other_inputs:
        for name in [n for n in ("julia", "random", "gradient") if n != args.input]:
            fill_input(name)
            for i in range(3):
                step(i)
            torch.cuda.synchronize()
            oev0, oev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_o = max(4, min(args.steps, 10))
            oev0.record(stream)
            for i in range(n_o):
                step(i)
            oev1.record(stream)
            torch.cuda.synchronize()
            o_ms = oev0.elapsed_time(oev1) / n_o
            other_inputs[name] = {"us_per_chain": 1e3 * o_ms, "GBps": chain_bytes / (o_ms * 1e-3) / 1e9,
                                  "frac_of_hbm_peak": None}
        if other_inputs:  # back to the headline input for the end-to-end leg
            fill_input(args.input)
            step(0), step(1)
            torch.cuda.synchronize()
    t_wall2 = time.time()

    # ---- BASELINE configs[4]: a batch of independent 4096^2 textures per rank (texture k -> rank k mod G, no
    # collective), nvpyrDispatchBatch = two launches for the whole batch ----
    batch = None
    if not args.no_batch:
        bw = bh = 4096
        per_rank = 32
        b_bytes = nv.chain_bytes(bw, bh)
        imgs = []
        for k in range(per_rank):
            t = torch.empty(b_bytes, dtype=torch.uint8, device=dev)
            t[:4 * bw * bh] = torch.randint(0, 256, (4 * bw * bh,), dtype=torch.uint8, device=dev, generator=gen)
            imgs.append(t)
        for _ in range(2):
            nv.dispatch_batch(stream, pipes, imgs, bw, bh)
        barrier()
        b_reps = 5
        bev0, bev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bev0.record(stream)
        for _ in range(b_reps):
            nv.dispatch_batch(stream, pipes, imgs, bw, bh)
        bev1.record(stream)
        barrier()
        b_ms = torch.tensor([bev0.elapsed_time(bev1) / b_reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(b_ms, op=dist.ReduceOp.MAX)
        b_ms = float(b_ms.item())
        batch = {"workload": f"{per_rank} independent 4096x4096 sRGBA8 textures per rank (uniform random bytes), one "
                             "nvpyrDispatchBatch per pass; 2.9 GB per rank per pass (> L2)",
                 "textures_per_s": world * per_rank / (b_ms * 1e-3), "us_per_texture_per_gpu": 1e3 * b_ms / per_rank,
                 "GBps": world * per_rank * b_bytes / (b_ms * 1e-3) / 1e9, "launches_per_batch": 2}
        del imgs

    # ---- end to end through the host-buffer entry point (nvpyrGenerateHost) ----
    # Headline: the reference's staging-buffer model (one host chain whose level 0 is filled, all other
    # levels filled on return -- scoped_image.hpp:436-453, and what cpuGenerateMipmaps_sRGBA does to a
    # MipmapStorage): H2D level 0, D2H levels 1..N-1.  Also timed: separate input/output buffers, where
    # level 0 travels back as well.  Upload, kernels and download are overlapped band by band inside the call.
    e2e = None
    if rank == 0 or world > 1:
        e_steps = max(2, min(args.steps, 5))
        l0_bytes = 4 * W * H
        h_in = torch.empty(l0_bytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(chain_bytes, dtype=torch.uint8).pin_memory()
        h_chain = torch.empty(chain_bytes, dtype=torch.uint8).pin_memory()
        h_in.copy_(bufs[0][:l0_bytes])
        h_chain[:l0_bytes].copy_(bufs[0][:l0_bytes])
        torch.cuda.synchronize()
        a_in, a_out, a_chain = h_in.numpy(), h_out.numpy(), h_chain.numpy()

        def timed(fn):
            fn()  # warm-up (allocates the library's device scratch)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                fn()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()) / e_steps

        t_inplace = timed(lambda: nv.generate_host(a_chain[:l0_bytes], W, H, out=a_chain))
        t_separate = timed(lambda: nv.generate_host(a_in, W, H, out=a_out))
        e2e = {"value": world * chain_bytes / t_inplace / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": l0_bytes, "d2h_bytes_per_step": chain_bytes - l0_bytes, "steps": e_steps,
               "ms_per_step": 1e3 * t_inplace, "host_cpus_bound_to_gpu_numa_node": numa_cpus,
               "api": "nvpyrGenerateHost in place on one pinned host chain (level 0 filled -> levels 1..14 filled), "
                      "upload / kernels / download overlapped in 32 MB bands",
               "separate_buffers": {"value": world * chain_bytes / t_separate / 1e9, "unit": UNIT,
                                    "ms_per_step": 1e3 * t_separate, "h2d_bytes_per_step": l0_bytes,
                                    "d2h_bytes_per_step": chain_bytes,
                                    "api": "nvpyrGenerateHost, pinned level 0 in, separate pinned packed chain out "
                                           "(level 0 downloaded too)"}}
        # sanity: both downloaded chains are what the device path produced, every byte
        want = bufs[0].cpu()
        assert torch.equal(h_out, want) and torch.equal(h_chain, want), "e2e chain differs from the device path"
        del want

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    sampler.stop()
    clocks = ClockSampler.summarise(sampler.window(t_wall0, t_wall2))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    for v in other_inputs.values():
        v["frac_of_hbm_peak"] = v["GBps"] / peak
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("fast6_srgba8_16384_bytes_per_launch")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "synthetic 16384x16384 sRGBA8 full mip chain, 15 levels (BASELINE configs[2]); level 0 = "
                               + {"julia": "the reference demo's Julia-set texture", "random": "uniform random bytes",
                                  "gradient": "smooth opaque gradient"}[args.input],
                   "algorithmic_bytes_per_step": chain_bytes, "us_per_chain": 1e3 * ms_per_step,
                   "l2_policy": "inputs larger than L2: two distinct 1.43 GB chains alternated",
                   "per_rank": "one chain per step per rank, no collective", "launches_per_chain": launches / args.steps,
                   "input": {"random": "uniform random bytes, all four channels (worst case for the encode table's bank "
                                       "conflicts)", "julia": "Julia set of the reference demo", "gradient":
                             "opaque smooth gradient"}[args.input],
                   "other_inputs": other_inputs, "batch_of_4096": batch},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "fastSrgba8Kernel<6> (TMA-staged level-0 slabs; level 0 -> levels 1..6)",
                     "algorithmic_bytes_per_launch": k_bytes, "us_per_launch": 1e3 * k_ms, "peak_source": peak_src},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-inputs", action="store_true")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--input", default="julia", choices=["random", "julia", "gradient"],
                    help="level-0 content of the headline loop. Default: the Julia-set texture the reference demo "
                         "regenerates and mip-maps every frame at its default size 16384x16384 "
                         "(demo_app/app_args.hpp:25, demo_app/julia.cpp:65-81). 'random' = uniform random bytes, the "
                         "worst case for the encode table's bank conflicts; every input is timed and reported.")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__), "--gpus",
               str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
