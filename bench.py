#!/usr/bin/env python
"""Benchmark of the mip-pyramid hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): full mip chain of a synthetic
16384x16384 sRGBA8 image, 15 levels, 1 431 655 764 algorithmic bytes (level 0 read once + every other level
written once).  A "step" is one full-chain generation.  The three level-0 contents of SURVEY 8d are each timed for
exactly K steps -- the Julia-set texture the reference demo mip-maps every frame at this very size, uniform random
bytes (worst case for shared-memory bank conflicts), a smooth gradient -- and reported under `inputs`; the headline
`value` / `ms_per_step` / `roofline` are those of the SLOWEST input.  Two distinct 1.43 GB chains are alternated, each
far larger than the 126 MB L2, so level 0 is never cache resident.  At N > 1 every rank runs the same step on its own
image (independent units, no data-path collective): weak scaling, value = N * bytes / max-over-ranks time.
`config.batch_of_4096` is BASELINE configs[4] as written: 512 textures of 4096^2 partitioned over the N ranks
(strong scaling), per-texture checksums gathered and compared with the single-GPU run.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU generator
(cpuGenerateMipmaps_sRGBA compiled in place from /root/reference into oracle/_ref, else our C port of it) on
the host cores, on the same 16384^2 configuration; that leg and the `cpu_baseline` object are the only places this
file touches oracle/.
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


def _stats(ms_list):
    """min / median / max per step (the reference's benchmark reports the same three, mipmaps_app.cpp:812-821)."""
    v = sorted(ms_list)
    return {"min_us": 1e3 * v[0], "median_us": 1e3 * v[len(v) // 2], "max_us": 1e3 * v[-1]}


def _host_threads():
    try:
        cores = len(os.sched_getaffinity(0))  # the host threads this process may actually use
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    return max(1, min(cores, 256))


def _julia_level0_host(w, h, threads):
    """Level 0 of the headline workload on the host: the Julia-set texture of the reference demo
    (shaders/julia.comp:27-63, demo_app/julia.cpp:65-81), filled by `threads` host threads."""
    import numpy as np
    so = os.path.join(ROOT, "oracle", "libnvpyr_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    lib = C.CDLL(so)
    lib.nvo_julia_srgba8_rows.argtypes = [C.c_void_p] + [C.c_uint32] * 5 + [C.c_int]
    out = np.empty(4 * w * h, dtype=np.uint8)
    rows = (h + threads - 1) // threads
    ts = [threading.Thread(target=lib.nvo_julia_srgba8_rows,
                           args=(out.ctypes.data, w, h, y, min(h, y + rows), 2109710467, 64))
          for y in range(0, h, rows)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


WORKLOAD = ("synthetic 16384x16384 sRGBA8 full mip chain, 15 levels, 1 431 655 764 algorithmic bytes "
            "(BASELINE configs[2])")


def cpu_baseline_single(sample_edge=16384):
    """The reference's CPU generator on ONE host thread (how the reference itself runs it, one std::thread
    per image, demo_app/mipmaps_app.cpp:651-652) over the whole workload of one step."""
    kind, make = _load_cpu_generator()
    l0 = _julia_level0_host(sample_edge, sample_edge, _host_threads())
    run, free = make(l0, sample_edge, sample_edge)
    t = time.perf_counter()
    run()
    dt = time.perf_counter() - t
    free()
    by = algorithmic_bytes(sample_edge, sample_edge)
    return {"value": by / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"one {sample_edge}x{sample_edge} sRGBA8 full chain (Julia-set level 0; the whole workload of one "
                      f"step), {dt:.2f} s on one host thread"}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU generator (cpuGenerateMipmaps_sRGBA, compiled in place into
    oracle/_ref) on the SAME configuration as the GPU arm: every step generates the full mip chain of the 16384^2
    Julia-set image, one chain per host thread on all host threads -- the reference's own parallelism (one
    std::thread per image, demo_app/mipmaps_app.cpp:651-652; generateLevel itself is single-threaded).
    value = threads * 1 431 655 764 B / step time.  About 11-15 s per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, make = _load_cpu_generator()
    cores = _host_threads()
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 64 << 30
    chain = algorithmic_bytes(W, H)
    threads = int(max(1, min(cores, (avail // 2) // chain)))
    l0 = _julia_level0_host(W, H, cores)
    jobs = [make(l0, W, H) for _ in range(threads)]
    del l0

    def step():
        ts = [threading.Thread(target=j[0]) for j in jobs]
        [t.start() for t in ts]
        [t.join() for t in ts]
    for _ in range(args.warmup):
        step()
    per_step = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t = time.perf_counter()
        step()
        per_step.append(1e3 * (time.perf_counter() - t))
    dt = time.perf_counter() - t0
    [j[1]() for j in jobs]
    value = chain * threads * args.steps / dt / 1e9
    sample = (f"each step = {threads} full 16384x16384 chains (the whole workload, Julia-set level 0), one per host "
              f"thread on {threads} of {cores} host threads: the reference's own thread-per-image parallelism")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "algorithmic_bytes_per_step": chain, "chains_per_step": threads,
                   "input": "Julia set of the reference demo", "per_step": _stats(per_step)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
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
def _cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads to the CPUs next to its GPU BEFORE any pinned host memory is allocated, so that
    the staging buffers of the end-to-end leg live on the GPU's own NUMA node (first touch): with one process per
    GPU the host side of the round trip otherwise crosses the socket interconnect for half of the ranks.  The node
    is taken from sysfs (/sys/bus/pci/devices/<bus id>/numa_node -> /sys/devices/system/node/nodeN/cpulist) and, where
    sysfs has no answer (-1: a VM without NUMA topology), from NVML's affinity mask.  Best effort: returns
    {"cpus": how many CPUs the rank is bound to, "node": sysfs node or None, "source": ...}."""
    info = {"cpus": 0, "node": None, "source": None}
    allowed = set(os.sched_getaffinity(0))
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["node"] = node
        if node >= 0:
            cpus = [c for c in _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) if c in allowed]
            if cpus:
                os.sched_setaffinity(0, cpus)
                info.update(cpus=len(cpus), source="sysfs")
                return info
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [64 * i + b for i, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(cpus=len(cpus), source="nvml")
    except Exception:
        pass
    return info


# Checksum of the 512 per-texture checksums of BASELINE configs[4] (texture k = uniform random bytes from seed
# BATCH_SEED + k), recorded from the single-GPU run: every sharded run must reproduce it (batch.fold_checksums).
BATCH_SEED = 100000
BATCH_EXPECTED = {512: 0x1ceab29ae8b61d5e}


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import vk_compute_mipmaps_b200 as nv
    from vk_compute_mipmaps_b200 import batch as nvbatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pipes = nv.PyramidPipelines()
    chain_bytes = nv.chain_bytes(W, H)
    assert chain_bytes == algorithmic_bytes(W, H) == 1431655764
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    bufs = [torch.empty(chain_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]  # two chains, alternated (each >> L2)

    def fill_input(name):
        if name == "random":
            for b in bufs:
                b[:4 * W * H] = torch.randint(0, 256, (4 * W * H,), dtype=torch.uint8, device=dev, generator=gen)
        else:
            {"julia": fill_julia, "gradient": fill_gradient}[name](bufs[0][:4 * W * H].view(H, W, 4), W, H)
            bufs[1][:4 * W * H].copy_(bufs[0][:4 * W * H])
    stream = torch.cuda.current_stream()
    warm = max(3, args.warmup)
    k_bytes = algorithmic_bytes(W, H, 0, 6)

    def timed_steps(fn, n_warm, n):
        """n_warm untimed + exactly n timed calls of fn(i), one event between consecutive steps; barrier +
        synchronize on both sides; returns (max-over-ranks total ms, this rank's per-step ms, launches)."""
        for i in range(n_warm):
            fn(i)
        barrier()
        l0 = nv.launch_count()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record(stream)
        for i in range(n):
            fn(i)
            evs[i + 1].record(stream)
        barrier()
        launches = nv.launch_count() - l0
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
        return max_over_ranks(evs[0].elapsed_time(evs[n])), per, launches

    # ---- the three synthetic level-0 contents of SURVEY 8d: whole chain, and the dominant kernel alone
    # (levelCount = 7 -> exactly the 6-level fast launch on level 0) ----
    names = [args.input] if args.no_other_inputs else ["julia", "random", "gradient"]
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    per_input = {}
    for name in names:
        fill_input(name)
        total_ms, per, launches = timed_steps(lambda i: nv.cmd_pyramid_dispatch(stream, pipes, W, H, image=bufs[i & 1]),
                                              warm, args.steps)
        k_total_ms, k_per, k_launches = timed_steps(
            lambda i: nv.cmd_pyramid_dispatch(stream, pipes, W, H, 7, image=bufs[i & 1]), 3, args.steps)
        assert k_launches == args.steps
        ms = total_ms / args.steps
        k_ms = k_total_ms / args.steps
        per_input[name] = {"ms_per_step": ms, "us_per_chain": 1e3 * ms, "GBps": world * chain_bytes / (ms * 1e-3) / 1e9,
                           "launches_per_chain": launches / args.steps, "launches": int(launches),
                           "per_step": _stats(per), "kernel_us": 1e3 * k_ms, "kernel_per_launch": _stats(k_per),
                           "kernel_GBps": k_bytes / (k_ms * 1e-3) / 1e9}
    t_wall2 = time.time()
    # headline = the WORST of the inputs (the bytes moved are identical; the encode table's bank conflicts are not)
    worst = min(per_input, key=lambda n: per_input[n]["GBps"])
    ms_per_step = per_input[worst]["ms_per_step"]
    value = per_input[worst]["GBps"]
    launches = per_input[worst]["launches"]
    if names[-1] != "julia":  # the end-to-end leg runs on the reference demo's texture
        fill_input("julia" if "julia" in names else args.input)
    nv.cmd_pyramid_dispatch(stream, pipes, W, H, image=bufs[0])
    torch.cuda.synchronize()

    # ---- BASELINE configs[4] as written: a batch of 512 independent 4096^2 textures PARTITIONED over the ranks
    # (texture k -> rank k % world, vk_compute_mipmaps_b200.batch.shard_indices; no data crosses GPUs), one
    # nvpyrDispatchBatch per rank and pass (two launches), per-texture checksums gathered and compared ----
    batch = None
    if not args.no_batch:
        bw = bh = 4096
        total_tex = args.batch_textures
        mine = nvbatch.shard_indices(total_tex, rank, world)
        b_bytes = nv.chain_bytes(bw, bh)
        stride = (b_bytes + 255) // 256 * 256
        pool = torch.empty(max(1, len(mine)) * stride, dtype=torch.uint8, device=dev)

        def texture_level0(k, out):
            g = torch.Generator(device=dev).manual_seed(BATCH_SEED + k)
            out[:4 * bw * bh] = torch.randint(0, 256, (4 * bw * bh,), dtype=torch.uint8, device=dev, generator=g)
        imgs = []
        for j, k in enumerate(mine):
            t = pool[j * stride:j * stride + b_bytes]
            texture_level0(k, t)
            imgs.append(t)
        b_reps = 3
        b_total_ms, b_per, b_launches = timed_steps(lambda i: nv.dispatch_batch(stream, pipes, imgs, bw, bh), 1, b_reps)
        b_ms = b_total_ms / b_reps
        # checksums: every rank fills its own slots, a SUM all_reduce assembles the table (no texture moves)
        sums = torch.zeros(total_tex, dtype=torch.int64, device=dev)
        for j, k in enumerate(mine):
            c = nvbatch.device_checksum(imgs[j])
            sums[k] = c - (1 << 64) if c >= (1 << 63) else c
        # cross-rank check on hardware: this rank also generates two textures OWNED BY THE NEXT RANK, alone
        # (one nvpyrDispatch each, not the fused batch), and their checksums must equal the owner's
        probe = nvbatch.shard_indices(total_tex, (rank + 1) % world, world)[:2] if total_tex else []
        probe_sums = torch.zeros(total_tex, dtype=torch.int64, device=dev)
        scratch = torch.empty(b_bytes, dtype=torch.uint8, device=dev)
        for k in probe:
            texture_level0(k, scratch)
            nv.cmd_pyramid_dispatch(stream, pipes, bw, bh, image=scratch)
            torch.cuda.synchronize()
            c = nvbatch.device_checksum(scratch)
            probe_sums[k] = c - (1 << 64) if c >= (1 << 63) else c
        probed = torch.zeros(total_tex, dtype=torch.int64, device=dev)
        if probe:
            probed[probe] = 1
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(probe_sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(probed, op=dist.ReduceOp.SUM)
        sel = probed > 0
        cross_ok = bool((sums[sel] == probe_sums[sel]).all().item()) and int(probed.max().item()) <= 1
        folded = nvbatch.fold_checksums([int(v) & 0xFFFFFFFFFFFFFFFF for v in sums.tolist()])
        expected = BATCH_EXPECTED.get(total_tex)
        batch = {"workload": f"{total_tex} independent 4096x4096 sRGBA8 textures (uniform random bytes, seed "
                             f"{BATCH_SEED}+k) partitioned over {world} rank(s) with batch.shard_indices, one "
                             "nvpyrDispatchBatch per rank and pass (> L2 per rank)",
                 "total_textures": total_tex, "textures_per_rank": len(mine), "scaling": "strong",
                 "textures_per_s": total_tex / (b_ms * 1e-3), "us_per_texture_per_gpu": 1e3 * b_ms / max(1, len(mine)),
                 "ms_per_pass": b_ms, "per_pass": _stats(b_per),
                 "GBps": total_tex * b_bytes / (b_ms * 1e-3) / 1e9, "launches_per_pass_per_rank": b_launches / b_reps,
                 "checksum_of_checksums": f"{folded:016x}",
                 "expected_from_single_gpu_run": None if expected is None else f"{expected:016x}",
                 "matches_single_gpu_run": None if expected is None else folded == expected,
                 "cross_rank_probes": int(sel.sum().item()), "cross_rank_probes_equal": cross_ok}
        assert cross_ok, "a texture generated alone on another rank differs from its owner's batch result"
        assert expected is None or folded == expected, "sharded batch differs from the single-GPU run"
        del imgs, pool, scratch

    # ---- end to end through the host-buffer entry point (nvpyrGenerateHost) ----
    # Headline: the reference's staging-buffer model (one host chain whose level 0 is filled, all other
    # levels filled on return -- scoped_image.hpp:436-453, and what cpuGenerateMipmaps_sRGBA does to a
    # MipmapStorage): H2D level 0, D2H levels 1..N-1.  Also timed: separate input/output buffers (level 0 travels
    # back as well), a caller with PAGEABLE memory (what examples/minimal_mipmaps.cpp passes: a std::vector), and
    # the bare H2D copy of level 0 (the floor of the round trip: it is PCIe-bound).
    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 5))
        l0_bytes = 4 * W * H
        h_in = torch.empty(l0_bytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(chain_bytes, dtype=torch.uint8).pin_memory()
        h_chain = torch.empty(chain_bytes, dtype=torch.uint8).pin_memory()
        h_in.copy_(bufs[0][:l0_bytes])
        h_chain[:l0_bytes].copy_(bufs[0][:l0_bytes])
        torch.cuda.synchronize()
        a_in, a_out, a_chain = h_in.numpy(), h_out.numpy(), h_chain.numpy()
        p_chain = np.empty(chain_bytes, dtype=np.uint8)  # pageable
        p_chain[:l0_bytes] = a_in

        def timed(fn):
            fn()  # warm-up (allocates the library's device scratch)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                fn()
            barrier()
            return max_over_ranks(time.perf_counter() - t0) / e_steps

        t_inplace = timed(lambda: nv.generate_host(a_chain[:l0_bytes], W, H, out=a_chain))
        t_separate = timed(lambda: nv.generate_host(a_in, W, H, out=a_out))
        t_pageable = timed(lambda: nv.generate_host(p_chain[:l0_bytes], W, H, out=p_chain))

        def h2d_only():
            bufs[1][:l0_bytes].copy_(h_in, non_blocking=True)
            torch.cuda.synchronize()
        t_h2d = timed(h2d_only)
        e2e = {"value": world * chain_bytes / t_inplace / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": l0_bytes, "d2h_bytes_per_step": chain_bytes - l0_bytes, "steps": e_steps,
               "ms_per_step": 1e3 * t_inplace, "host_cpus_bound_to_gpu_numa_node": numa["cpus"],
               "gpu_numa_node_sysfs": numa["node"], "numa_binding_source": numa["source"],
               "api": "nvpyrGenerateHost in place on one pinned host chain (level 0 filled -> levels 1..14 filled), "
                      "upload / kernels / download overlapped in 32 MB bands",
               "separate_buffers": {"value": world * chain_bytes / t_separate / 1e9, "unit": UNIT,
                                    "ms_per_step": 1e3 * t_separate, "h2d_bytes_per_step": l0_bytes,
                                    "d2h_bytes_per_step": chain_bytes,
                                    "api": "nvpyrGenerateHost, pinned level 0 in, separate pinned packed chain out "
                                           "(level 0 downloaded too)"},
               "pageable_caller": {"value": world * chain_bytes / t_pageable / 1e9, "unit": UNIT,
                                   "ms_per_step": 1e3 * t_pageable,
                                   "api": "nvpyrGenerateHost in place on a PAGEABLE host chain (malloc'd memory, what "
                                          "examples/minimal_mipmaps.cpp passes)"},
               "h2d_only": {"GBps_per_rank": l0_bytes / t_h2d / 1e9, "GBps_all_ranks": world * l0_bytes / t_h2d / 1e9,
                            "ms": 1e3 * t_h2d, "what": "cudaMemcpyAsync of level 0 alone, pinned -> device, all ranks at "
                                                       "once: the floor of the round trip"}}
        # sanity: the downloaded chains are what the device path produced, every byte
        want = bufs[0].cpu()
        assert torch.equal(h_out, want) and torch.equal(h_chain, want) and bool((torch.from_numpy(p_chain) == want).all()), \
            "e2e chain differs from the device path"
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
    for v in per_input.values():
        v["frac_of_hbm_peak"] = v["GBps"] / world / peak
        v["kernel_frac_of_hbm_peak"] = v["kernel_GBps"] / peak
    achieved = per_input[worst]["kernel_GBps"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("fast6_srgba8_16384_bytes_per_launch")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single()
    described = {"random": "uniform random bytes, all four channels (worst case for the encode table's bank conflicts)",
                 "julia": "Julia set of the reference demo (the texture it mip-maps every frame at this size)",
                 "gradient": "opaque smooth gradient"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "algorithmic_bytes_per_step": chain_bytes, "us_per_chain": 1e3 * ms_per_step,
                   "input": described[worst] + " -- the SLOWEST of the inputs timed; all of them under `inputs`",
                   "l2_policy": "inputs larger than L2: two distinct 1.43 GB chains alternated",
                   "per_rank": "one chain per step per rank, no collective",
                   "launches_per_chain": per_input[worst]["launches_per_chain"], "per_step": per_input[worst]["per_step"],
                   "batch_of_4096": batch},
        "inputs": {n: dict(v, description=described[n]) for n, v in per_input.items()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "fastSrgba8Kernel<6> (TMA-staged level-0 slabs, lane-private decode and encode tables; level 0 -> levels 1..6)",
                     "input": worst, "algorithmic_bytes_per_launch": k_bytes, "us_per_launch": per_input[worst]["kernel_us"],
                     "per_launch": per_input[worst]["kernel_per_launch"], "peak_source": peak_src,
                     "per_input": {n: {"us_per_launch": v["kernel_us"], "frac": v["kernel_frac_of_hbm_peak"]}
                                   for n, v in per_input.items()}},
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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--batch-textures", type=int, default=512,
                    help="BASELINE configs[4]: total textures of the 4096^2 batch, partitioned over the ranks")
    ap.add_argument("--input", default="julia", choices=["random", "julia", "gradient"],
                    help="with --no-other-inputs: the only level-0 content timed (default: the Julia-set texture of "
                         "the reference demo, demo_app/app_args.hpp:25, demo_app/julia.cpp:65-81). Without it all "
                         "three contents are timed and the slowest one is the headline.")
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
