O=gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/r2w_gputests.txt 2>&1
cat $O/r2w_gputests.txt
for rep in 1 2; do for lib in libnvpyr.so libnvpyr_prev.so; do echo "== $lib"; LD_PRELOAD=$PWD/vk_compute_mipmaps_b200/$lib tools/bench_native --batches 30 2>&1 | grep -E "409[45]|2047|2052|mandel|1080p|tall"; done; done
