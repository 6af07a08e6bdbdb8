"""Times every BASELINE.json config class on one GPU (not the headline bench; see bench.py).
Protocol of the reference's benchmark (demo_app/mipmaps_app.cpp:621-622,712-728): batches of 8 back-to-back
generations between two events, first batch discarded, min / median per generation.
usage: python tools/bench_configs.py [--batches 40] [--out profiles/xxx.json]"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vk_compute_mipmaps_b200 as nv

CONFIGS = [  # name, w, h, fmt, RTX 3090 min ns of the reference (README.md:166-189), or None
    ("4096.jpg class", 4096, 4096, 0, 112000), ("4095.jpg class", 4095, 4095, 0, 188032),
    ("4094.jpg class", 4094, 4094, 0, 161664), ("lunch_2047 class", 2047, 2047, 0, 71592),
    ("2048 class", 2048, 2048, 0, 35712), ("1080p class", 1920, 1080, 0, 36736),
    ("1440p class", 2560, 1440, 0, 43008), ("4k class", 3840, 2160, 0, 75520),
    ("tall class", 1080, 4096, 0, 50952), ("alpha2052 class", 2052, 2052, 0, 44928),
    ("mandelbrots class", 3095, 990, 0, 57472), ("16384 synthetic", 16384, 16384, 0, None),
    ("8192 synthetic", 8192, 8192, 0, None), ("4096 rgba32f", 4096, 4096, 1, None),
    ("4095 rgba32f", 4095, 4095, 1, None),
]

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--batches", type=int, default=30); ap.add_argument("--out")
    ap.add_argument("--only", default="")
    ap.add_argument("--graph", action="store_true", help="also time the chain captured once into a CUDA graph and "
                    "replayed (the reference records a command buffer once and submits it per frame)")
    a = ap.parse_args()
    peak = 6541.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p): peak = float(json.load(open(p))["hbm_gbs"])
    res = []
    for name, w, h, fmt, ref_ns in CONFIGS:
        if a.only and a.only not in name: continue
        pipes = nv.PyramidPipelines(format=fmt)
        nbytes = nv.chain_bytes(w, h, 0, fmt)
        nrot = max(2, min(8, int(400e6 // nbytes) + 1))  # rotate buffers so that small chains are not L2 resident
        bufs = []
        for k in range(nrot):
            if fmt == 0:
                b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
                b[:4 * w * h] = torch.randint(0, 256, (4 * w * h,), dtype=torch.uint8, device="cuda")
            else:
                b = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
                b[:4 * w * h] = torch.rand(4 * w * h, device="cuda")
            bufs.append(b)
        st = torch.cuda.current_stream()
        l0 = nv.launch_count()
        nv.cmd_pyramid_dispatch(st, pipes, w, h, image=bufs[0])
        launches = nv.launch_count() - l0
        times = []
        for bi in range(a.batches + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(8):
                nv.cmd_pyramid_dispatch(st, pipes, w, h, image=bufs[(bi * 8 + i) % nrot])
            e1.record(st)
            torch.cuda.synchronize()
            if bi: times.append(e0.elapsed_time(e1) * 1e6 / 8)
        times.sort()
        mn, med = times[0], times[len(times) // 2]
        graph_med = None
        if a.graph:
            side = torch.cuda.Stream()
            graphs = []
            for k in range(nrot):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    nv.cmd_pyramid_dispatch(None, pipes, w, h, image=bufs[k])
                graphs.append(g)
            gt = []
            for bi in range(a.batches + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for i in range(8):
                    graphs[(bi * 8 + i) % nrot].replay()
                e1.record(st)
                torch.cuda.synchronize()
                if bi: gt.append(e0.elapsed_time(e1) * 1e6 / 8)
            gt.sort()
            graph_med = gt[len(gt) // 2]
        r = {"config": name, "w": w, "h": h, "format": "srgba8" if fmt == 0 else "rgba32f", "launches": launches,
             "min_ns": round(mn), "median_ns": round(med), "algorithmic_bytes": nbytes,
             "GBps_at_median": round(nbytes / med, 1), "frac_of_hbm_peak": round(nbytes / med / peak, 3),
             "rtx3090_reference_min_ns": ref_ns, "rotating_buffers": nrot}
        if graph_med is not None:
            r["graph_replay_median_ns"] = round(graph_med)
        res.append(r)
        print(f"{name:20s} {w}x{h} {r['format']:8s} launches {launches:2d}  min {mn/1e3:9.1f} us  median {med/1e3:9.1f} us  "
              f"{r['GBps_at_median']:8.1f} GB/s  ({100*r['frac_of_hbm_peak']:.1f}% of HBM peak)"
              + (f"  [graph replay {graph_med/1e3:.1f} us]" if graph_med is not None else "")
              + (f"  [RTX3090 ref {ref_ns/1e3:.1f} us]" if ref_ns else ""), flush=True)
        del bufs
        torch.cuda.empty_cache()
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)

if __name__ == "__main__":
    main()
