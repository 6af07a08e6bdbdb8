"""BASELINE configs[4]: a batch of independent 4096^2 sRGBA8 textures on one GPU (the multi-GPU run shards
the batch, texture k -> rank k mod G, no collective).  Times nvpyrDispatchBatch (fused: two launches per batch)
against one dispatch per texture.  usage: python tools/bench_batch.py [--textures 64] [--size 4096] [--reps 5]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vk_compute_mipmaps_b200 as nv

ap = argparse.ArgumentParser()
ap.add_argument("--textures", type=int, default=64); ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--out")
a = ap.parse_args()
w = h = a.size
n = nv.chain_bytes(w, h)
peak = 6541.8
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk): peak = float(json.load(open(pk))["hbm_gbs"])
imgs = []
for k in range(a.textures):
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    b[:4 * w * h] = torch.randint(0, 256, (4 * w * h,), dtype=torch.uint8, device="cuda")
    imgs.append(b)
pipes, st = nv.PyramidPipelines(), torch.cuda.current_stream()
def fused(): nv.dispatch_batch(st, pipes, imgs, w, h)
def loop():
    for b in imgs: nv.cmd_pyramid_dispatch(st, pipes, w, h, image=b)
res = {"textures": a.textures, "size": [w, h], "bytes_per_texture": n, "total_GB": a.textures * n / 1e9}
for name, fn in (("fused_batch", fused), ("per_texture", loop)):
    fn(); torch.cuda.synchronize()
    l0 = nv.launch_count(); fn(); launches = nv.launch_count() - l0
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    res[name] = {"ms": ms, "us_per_texture": 1e3 * ms / a.textures, "GBps": a.textures * n / ms / 1e6,
                 "frac_of_hbm_peak": a.textures * n / ms / 1e6 / peak, "launches": launches}
    print(f"{name:12s} {a.textures} x {w}^2: {ms:8.3f} ms  {1e3 * ms / a.textures:7.2f} us/texture  "
          f"{a.textures * n / ms / 1e6:7.1f} GB/s ({100 * a.textures * n / ms / 1e6 / peak:.1f}% of HBM peak)  {launches} launches")
if a.out: json.dump(res, open(a.out, "w"), indent=1)
