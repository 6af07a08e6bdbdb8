"""Premultiply-alpha + chain at 16384^2: fused into the level-0 read (NVPYR_FLAG_PREMULTIPLY_ALPHA) against the
stand-alone pre-pass followed by the chain.  usage: python tools/bench_premultiply.py [--size 16384]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vk_compute_mipmaps_b200 as nv
ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=16384); ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--alpha", default="random", choices=["random", "sprite"],
                help="random: every texel has a random alpha (worst case); sprite: opaque except a 10 %% band of "
                     "random alpha and a 20 %% fully transparent band (typical cut-out texture)")
a = ap.parse_args()
w = h = a.size
n, l0 = nv.chain_bytes(w, h), 4 * w * h
src = torch.randint(0, 256, (l0,), dtype=torch.uint8, device="cuda")
if a.alpha == "sprite":
    v = src.view(h, w, 4)
    v[..., 3] = 255
    v[h // 2:h // 2 + h // 10, :, 3] = torch.randint(0, 256, (h // 10, w), dtype=torch.uint8, device="cuda")
    v[:h // 5, :, 3] = 0
bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
st, pipes = torch.cuda.current_stream(), nv.PyramidPipelines()
def fused(b): nv.cmd_pyramid_dispatch(st, pipes, w, h, image=b, flags=nv.FLAG_PREMULTIPLY_ALPHA)
def separate(b):
    nv.premultiply_alpha(st, b, b, w * h)
    nv.cmd_pyramid_dispatch(st, pipes, w, h, image=b)
for name, fn in (("fused", fused), ("separate", separate), ("no premultiply", lambda b: nv.cmd_pyramid_dispatch(st, pipes, w, h, image=b))):
    ts = []
    for r in range(a.reps + 1):
        b = bufs[r & 1]
        b[:l0] = src                      # level 0 is rewritten in place: restore the straight-alpha image
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(b); e1.record(st); torch.cuda.synchronize()
        if r: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:16s} {w}x{h} {a.alpha} alpha: median {ts[len(ts) // 2]:8.1f} us  min {ts[0]:8.1f} us")
