O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/t_default.log 2>&1; tail -3 $O/t_default.log
bash tools/ab_libs.sh random libnvpyr.so libnvpyr_w8s16.so 2>&1 | grep chain_us | tee $O/ab_random.txt
bash tools/ab_libs.sh julia libnvpyr.so libnvpyr_w8s16.so 2>&1 | grep chain_us | tee $O/ab_julia.txt
python tools/warm_launches.py --only "4096.jpg,4095.jpg,2047,1080p,2048,16384,alpha2052" --sizes 256x256,255x255,64x64,63x63 2>&1 | tee $O/warm.txt
