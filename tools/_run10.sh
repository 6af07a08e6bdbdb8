O=gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $O/r2j_gputests.txt 2>&1
cat $O/r2j_gputests.txt
python tools/bench_configs.py --batches 30 --out $O/r2j_configs.json 2>&1 | grep -v "^$" | tee $O/r2j_configs.txt
python tools/bench_batch.py --textures 128 --out $O/r2j_batch128.json 2>&1 | tail -5
