"""Warm per-kernel durations of every launch of a config's chain (CUPTI via torch.profiler; no cache flush, no
serialisation -- unlike the ncu launch list): python tools/warm_launches.py [--only substr] [--reps 20]"""
import argparse, collections, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vk_compute_mipmaps_b200 as nv
from bench_configs import CONFIGS

ap = argparse.ArgumentParser(); ap.add_argument("--only", default=""); ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--sizes", default="", help="extra sRGBA8 sizes, e.g. 256x256,63x63")
a = ap.parse_args()
CONFIGS = list(CONFIGS) + [(f"size {s}", int(s.split("x")[0]), int(s.split("x")[1]), 0, None) for s in a.sizes.split(",") if s]
for name, w, h, fmt, _ in CONFIGS:
    if a.only and not any(o in name for o in a.only.split(",")) and not name.startswith("size "): continue
    n = nv.chain_bytes(w, h, 0, fmt)
    nrot = max(2, min(8, int(400e6 // n) + 1))
    bufs = [torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda") if fmt == 0 else torch.rand(n // 4, device="cuda")
            for _ in range(nrot)]
    pipes = nv.PyramidPipelines(format=fmt)
    for i in range(4):
        nv.cmd_pyramid_dispatch(None, pipes, w, h, image=bufs[i % nrot])
    torch.cuda.synchronize()
    l0 = nv.launch_count()
    nv.cmd_pyramid_dispatch(None, pipes, w, h, image=bufs[0])
    per_chain = nv.launch_count() - l0
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(a.reps):
            nv.cmd_pyramid_dispatch(None, pipes, w, h, image=bufs[i % nrot])
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "ernel" in e.name),
                 key=lambda e: e.time_range.start)
    assert len(evs) == a.reps * per_chain, (len(evs), a.reps, per_chain)
    print(f"{name}  {w}x{h}: {per_chain} launches per chain")
    tot = 0.0
    for k in range(per_chain):
        d = sorted(evs[r * per_chain + k].time_range.elapsed_us() for r in range(a.reps))
        gaps = sorted(evs[r * per_chain + k].time_range.start - evs[r * per_chain + k - 1].time_range.end
                      for r in range(a.reps) if r * per_chain + k > 0)
        nm = evs[k].name.replace("void ", "").replace("nvpyr::", "")[:60]
        print(f"   {d[len(d) // 2]:8.2f} us (min {d[0]:7.2f})  gap before {gaps[len(gaps) // 2]:6.2f} us  {nm}")
        tot += d[len(d) // 2] + gaps[len(gaps) // 2]
    print(f"   sum incl. gaps {tot:8.2f} us")
