O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/t_default.log 2>&1; tail -4 $O/t_default.log | head -2
for rep in 1 2; do
for lib in libnvpyr.so libnvpyr_g1.so; do echo "== $lib"; NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/$lib python tools/bench_configs.py --only "class" --batches 20 2>&1 | grep -E "4095|4094|2047|2052|1080p|3095"; done; done
echo "== strip4 threshold 0"; NVPYR_GEN_STRIP4_MIN_TEXELS=0 python tools/bench_configs.py --only "class" --batches 20 2>&1 | grep -E "4095|2047|2052|1080p|3095|tall"
echo "== strip4 threshold 2^20"; NVPYR_GEN_STRIP4_MIN_TEXELS=1048576 python tools/bench_configs.py --only "class" --batches 20 2>&1 | grep -E "4095|2047|2052|1080p|3095|tall"
echo "== tail max 256^2"; NVPYR_TAIL_MAX_TEXELS=65536 python tools/bench_configs.py --only "class" --batches 20 2>&1 | grep -E "4095|2047|2052|1080p|3095|tall|4096|2048"
echo "== tail max 1024^2"; NVPYR_TAIL_MAX_TEXELS=1048576 python tools/bench_configs.py --only "class" --batches 20 2>&1 | grep -E "4095|2047|2052|1080p|3095|tall|4096|2048"
