O=gpurun_out; mkdir -p $O
./examples/custom_functors > $O/custom.txt 2>&1; echo "custom rc=$?"; tail -12 $O/custom.txt
( time python -m pytest tests -m gpu -x -q ) > $O/t_default.log 2>&1; tail -4 $O/t_default.log
