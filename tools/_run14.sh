O=gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/r2n_gputests.txt 2>&1
cat $O/r2n_gputests.txt
./examples/custom_functors | grep -i "MinFirst\|ok\|FAILED" | tail -25
