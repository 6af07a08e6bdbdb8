O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $O/r2k_gputests.txt 2>&1
cat $O/r2k_gputests.txt
for e in 0 1; do echo "== NVPYR_NO_SOLO_SMEM=$e"; NVPYR_NO_SOLO_SMEM=$e python tools/bench_configs.py --batches 20 2>&1 | grep -v "^$\|synthetic"; done | tee $O/r2k_cfg.txt
python tools/warm_launches.py --only "4095.jpg,lunch,1080p,4096.jpg,2052,mandel" 2>&1 | grep -v "^$\|Warn\|warn" | tee $O/r2k_warm.txt
