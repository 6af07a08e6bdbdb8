O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/t_default.log 2>&1; tail -3 $O/t_default.log
L="libnvpyr_w8s16.so libnvpyr_u1c.so libnvpyr_u1.so libnvpyr.so libnvpyr_u2np.so"
bash tools/ab_libs.sh random $L 2>&1 | grep chain_us | tee $O/ab_random.txt
bash tools/ab_libs.sh julia $L 2>&1 | grep chain_us | tee $O/ab_julia.txt
