O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "recorded_once or user_defined" 2>&1 | tail -15
python tools/bench_configs.py --batches 20 --graph --out $O/configs_graph.json 2>&1 | tee $O/configs_graph.txt | tail -20
