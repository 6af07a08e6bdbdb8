O=gpurun_out
for lib in libnvpyr.so libnvpyr_w24.so; do echo "== $lib"; NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python tools/warm_launches.py --only "4096.jpg,8192,2048,1080p,16384,4095.jpg,lunch" 2>&1 | grep -v "^$"; done | tee $O/r2e_warm.txt
