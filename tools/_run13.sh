O=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:generalStrip4Kernel --launch-skip 2 -c 1 -f -o $O/r2m_gen4095 python tools/launch_probe.py --only 4095.jpg --reps 4 > $O/r2m_ncu.log 2>&1
tail -2 $O/r2m_ncu.log
