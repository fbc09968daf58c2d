cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_dp.py tests/test_gpu_train.py -m gpu -q -x -k "dp or graph or replica" 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2f_bench_2gpu.json | head -c 600; echo
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
PY
