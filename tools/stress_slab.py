"""Stress of the slab-task hand-off of the tuned fast kernel (store -> fence -> shared atomic -> load, a pattern
compute-sanitizer's racecheck cannot vouch for): many repetitions on varied sizes, every result compared on the device
with the first one, and a checksum of the first printed for comparison across modes.
usage: python tools/stress_slab.py [--reps 300] [--sizes 1024x1024,...]
Run with NVPYR_SLAB_MAX_TILES_PER_WARP_X100=1000000 to force slab tasks (and stash-slot recycling) onto every size,
with NVPYR_NO_SLAB_TASKS=1 for the tile-mode answer."""
import argparse, hashlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vk_compute_mipmaps_b200 as nv

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=300)
ap.add_argument("--sizes", default="1024x1024,2048x2048,512x2048,1536x1024,64x4096,4096x4096,8192x8192")
a = ap.parse_args()
bad = 0
for s in a.sizes.split(","):
    w, h = (int(v) for v in s.split("x"))
    n = nv.chain_bytes(w, h, 0, 0)
    g = torch.Generator(device="cuda").manual_seed(w * 7 + h)
    l0 = torch.randint(0, 256, (4 * w * h,), dtype=torch.uint8, device="cuda", generator=g)
    reps = max(4, min(a.reps, int(a.reps * (1 << 22) / (w * h))))
    ref, diff = None, 0
    bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
    for r in range(reps):
        b = bufs[r & 1]
        b[4 * w * h:].fill_(0xAB)
        b[:4 * w * h] = l0
        nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=b)
        if ref is None:
            torch.cuda.synchronize()
            ref = b.clone()
        else:
            diff += int(not torch.equal(b, ref))
    torch.cuda.synchronize()
    bad += diff
    print(f"{w}x{h} reps {reps} differing {diff} sha256 {hashlib.sha256(ref.cpu().numpy().tobytes()).hexdigest()[:16]}", flush=True)
sys.exit(1 if bad else 0)
