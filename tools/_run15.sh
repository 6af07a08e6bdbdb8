O=gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_v3_bench_${N}gpus.json 2> $O/r2_v3_bench_${N}gpus.err
tail -c 600 $O/r2_v3_bench_${N}gpus.err
python - <<PY
import json
d=json.loads(open('$O/r2_v3_bench_${N}gpus.json').read().strip().splitlines()[-1])
print("N", d['n_gpus'], "value", round(d['value'],1), "ms/step", round(d['ms_per_step'],4))
b=d['config']['batch_of_4096']; print("batch", {k:b[k] for k in ('total_textures','textures_per_rank','textures_per_s','us_per_texture_per_gpu','matches_single_gpu_run','cross_rank_probes_equal')})
e=d['e2e']; print("e2e", round(e['value'],1), round(e['ms_per_step'],2), "sep", round(e['separate_buffers']['ms_per_step'],2), "pageable", round(e['pageable_caller']['ms_per_step'],2), "h2d_only", e['h2d_only'], e.get('gpu_numa_node_sysfs'), e.get('numa_binding_source'), e.get('host_cpus_bound_to_gpu_numa_node'))
PY
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" | head -10
