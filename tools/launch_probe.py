"""Runs each config class a few times so that `ncu --metrics gpu__time_duration.sum` lists every launch of
its chain: python tools/launch_probe.py [--only substr] [--reps 3]   (run it UNDER ncu; prints nothing timed)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vk_compute_mipmaps_b200 as nv
from bench_configs import CONFIGS

ap = argparse.ArgumentParser(); ap.add_argument("--only", default=""); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
for name, w, h, fmt, _ in CONFIGS:
    if a.only and a.only not in name: continue
    n = nv.chain_bytes(w, h, 0, fmt)
    if fmt == 0:
        b = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
    else:
        b = torch.rand(n // 4, device="cuda")
    torch.cuda.synchronize()
    for _ in range(a.reps):
        nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(format=fmt), w, h, image=b)
    torch.cuda.synchronize()
    print(name, w, h, flush=True)
