O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $O/r2b_gputests.txt 2>&1
cat $O/r2b_gputests.txt
bash tools/ab_inputs.sh libnvpyr.so libnvpyr_prev.so libnvpyr.so libnvpyr_prev.so 2>&1 | tee $O/r2b_ab_inputs.txt
bash tools/ab_cfg.sh "" libnvpyr.so libnvpyr_prev.so 2>&1 | tee $O/r2b_ab_cfg.txt
