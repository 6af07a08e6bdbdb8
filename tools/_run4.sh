O=gpurun_out
bash tools/ab_inputs.sh libnvpyr.so libnvpyr_w24.so libnvpyr_w16.so libnvpyr.so libnvpyr_w24.so libnvpyr_w16.so 2>&1 | tee $O/r2d_ab_inputs.txt
bash tools/ab_cfg.sh "" libnvpyr_w24.so 2>&1 | tee $O/r2d_ab_cfg.txt
