set -x
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $O/r2a_gputests.txt 2>&1
tail -5 $O/r2a_gputests.txt
(time python bench.py --steps 20 --warmup 3) > $O/r2a_bench.json 2> $O/r2a_bench.err
tail -c 3000 $O/r2a_bench.json
python tools/bench_configs.py --batches 30 --out $O/r2a_configs.json > $O/r2a_configs.txt 2>&1
cat $O/r2a_configs.txt
