O=gpurun_out
for v in "32:libnvpyr.so" "0:libnvpyr.so" "0:libnvpyr_w28.so" "0:libnvpyr_w20.so"; do
  e=${v%%:*}; lib=${v##*:}; echo "== $lib NVPYR_FAST_WARPS_LARGE=$e"
  NVPYR_FAST_WARPS_LARGE=$e NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python tools/bench_configs.py --only "synthetic" --batches 20 2>&1 | grep -v "^$"
  NVPYR_FAST_WARPS_LARGE=$e NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python tools/bench_configs.py --only "4096.jpg" --batches 20 2>&1 | grep -v "^$"
  NVPYR_FAST_WARPS_LARGE=$e NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python tools/bench_configs.py --only "4k" --batches 20 2>&1 | grep -v "^$"
done 2>&1 | tee $O/r2g_warps.txt
