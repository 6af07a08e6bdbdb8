O=gpurun_out
for e in 0 1024 4096; do echo "== NVPYR_SOLO_SMEM_MAX_TEXELS=$e"; NVPYR_SOLO_SMEM_MAX_TEXELS=$e python tools/bench_configs.py --batches 20 2>&1 | grep -v "^$\|synthetic\|rgba32f"; done | tee $O/r2l_cfg.txt
python tools/warm_launches.py --only "4095.jpg,lunch,1080p" 2>&1 | grep -v "^$\|Warn\|warn" | tee $O/r2l_warm.txt
