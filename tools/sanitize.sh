#!/bin/bash
# compute-sanitizer over a small but representative set of dispatches (all kernels, all schedule classes).
# usage (GPU box): tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]
TOOL=${1:-memcheck}
SLAB=${2:-25}   # second argument 1000000: slab tasks (and stash-slot recycling) on every size in the second pass
cat > /tmp/san_driver.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import vk_compute_mipmaps_b200 as nv, _oracle
o = _oracle.load_oracle()
cases = [(256, 256, 0, False), (1024, 512, 0, False), (255, 255, 0, False), (260, 260, 0, False), (136, 512, 0, False),
         (777, 1031, 0, False), (1200, 900, 0, False), (64, 64, 1, False), (100, 37, 1, False), (96, 160, 0, True),
         (2560, 2048, 0, False)]  # the last one: 1280 tiles = the TMA-staged tile mode of the fast kernel
for (w, h, fmt, fg) in cases:
    l0 = _oracle.random_level0(w, h, 5, fmt=fmt)
    dt = torch.uint8 if fmt == 0 else torch.float32
    buf = torch.zeros(nv.chain_bytes(w, h, 0, fmt) // (1 if fmt == 0 else 4), dtype=dt, device='cuda')
    buf[:4 * w * h] = torch.from_numpy(l0).cuda()
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(format=fmt, fast_pipeline=not fg), w, h, image=buf)
    torch.cuda.synchronize()
    want = o.shader_chain(l0, w, h, fmt=fmt, force_general=fg)[0]
    got = buf.cpu().numpy()
    assert (got.view(np.uint8) == want.view(np.uint8)).all(), (w, h, fmt, fg)
# fused batch (+ premultiply in the first launch), fused premultiply, stand-alone premultiply, banded host pipeline
w, h = 1024, 512
imgs, wants = [], []
for k in range(3):
    l0 = _oracle.random_level0(w, h, 40 + k)
    wants.append(o.shader_chain(o.premultiply(l0), w, h)[0])
    b = torch.zeros(nv.chain_bytes(w, h), dtype=torch.uint8, device='cuda'); b[:4 * w * h] = torch.from_numpy(l0).cuda(); imgs.append(b)
nv.dispatch_batch(None, nv.PyramidPipelines(), imgs, w, h, flags=nv.FLAG_PREMULTIPLY_ALPHA)
torch.cuda.synchronize()
assert all((b.cpu().numpy() == x).all() for b, x in zip(imgs, wants)), 'batch'
for (w, h) in [(1024, 768), (1023, 300)]:
    l0 = _oracle.random_level0(w, h, 50)
    b = torch.zeros(nv.chain_bytes(w, h), dtype=torch.uint8, device='cuda'); b[:4 * w * h] = torch.from_numpy(l0).cuda()
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=b, flags=nv.FLAG_PREMULTIPLY_ALPHA)
    torch.cuda.synchronize()
    assert (b.cpu().numpy() == o.shader_chain(o.premultiply(l0), w, h)[0]).all(), ('premultiply', w, h)
for (w, h, fmt) in [(333, 97, 0), (260, 260, 0), (100, 37, 1)]:  # the blit fallback (NVPYR_FLAG_GENERAL_BLIT)
    l0 = _oracle.random_level0(w, h, 70, fmt=fmt)
    dt = torch.uint8 if fmt == 0 else torch.float32
    b = torch.zeros(nv.chain_bytes(w, h, 0, fmt) // (1 if fmt == 0 else 4), dtype=dt, device='cuda'); b[:4 * w * h] = torch.from_numpy(l0).cuda()
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(format=fmt), w, h, image=b, flags=nv.FLAG_GENERAL_BLIT)
    torch.cuda.synchronize()
    assert (b.cpu().numpy().view(np.uint8) == o.shader_chain(l0, w, h, fmt=fmt, general_blit=True)[0].view(np.uint8)).all(), ('blit', w, h)
w, h = 512, 1088
l0 = _oracle.random_level0(w, h, 60)
assert (nv.generate_host(l0, w, h) == o.shader_chain(l0, w, h)[0]).all(), 'host pipeline'
print('sanitize driver ok')
PY
for tail in 262144 0; do
  echo "== $TOOL NVPYR_TAIL_MAX_TEXELS=$tail (second pass: 24-warp build of the fast kernel, slab tasks forced onto every size)"
  [ $tail = 0 ] && export NVPYR_FAST_WARPS_LARGE=24 NVPYR_SLAB_MAX_TILES_PER_WARP_X100=$SLAB
  NVPYR_HOST_BAND_BYTES=131072 NVPYR_TAIL_MAX_TEXELS=$tail compute-sanitizer --tool $TOOL --kernel-regex kns=nvpyr --print-limit 20 python /tmp/san_driver.py 2>&1 | tail -6
done
echo "== $TOOL NVPYR_CASCADE=1 (general dispatches of the small levels as cascades: cascadeRun, off by default)"
NVPYR_CASCADE=1 NVPYR_HOST_BAND_BYTES=131072 compute-sanitizer --tool $TOOL --kernel-regex kns=nvpyr --print-limit 20 python /tmp/san_driver.py 2>&1 | tail -6
echo "== $TOOL examples/custom_functors (user-defined functor sets through include/nvpyr.cuh)"
compute-sanitizer --tool $TOOL --print-limit 20 ./examples/custom_functors 2>&1 | tail -4
