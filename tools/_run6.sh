O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $O/r2f_gputests.txt 2>&1
cat $O/r2f_gputests.txt
bash tools/ab_inputs.sh libnvpyr.so libnvpyr_w24.so libnvpyr.so libnvpyr_w24.so 2>&1 | tee $O/r2f_ab_inputs.txt
NVPYR_FAST_WARPS_LARGE=32 bash tools/ab_inputs.sh libnvpyr.so 2>&1 | tee -a $O/r2f_ab_inputs.txt
bash tools/ab_cfg.sh "" libnvpyr.so 2>&1 | tee $O/r2f_ab_cfg.txt
