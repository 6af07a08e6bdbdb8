set -x
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/t_default.log 2>&1; tail -3 $O/t_default.log
for v in w16c w8c w4s16; do NVPYR_LIB_PATH=$PWD/vk_compute_mipmaps_b200/libnvpyr_$v.so python -m pytest tests -m gpu -q -x > $O/t_$v.log 2>&1; tail -2 $O/t_$v.log; done
L="libnvpyr_w4s15.so libnvpyr_w8s16.so libnvpyr_w16c.so libnvpyr_w8c.so libnvpyr_w4s16.so libnvpyr_w4s15pad.so"
bash tools/ab_libs.sh random $L 2>&1 | grep chain_us | tee $O/ab_random.txt
bash tools/ab_libs.sh julia $L 2>&1 | grep chain_us | tee $O/ab_julia.txt
