cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c6_base$i.json 2>> gpurun_out/c6.err; echo "rc=$?"
done
timeout 120 python bench.py --workload rawiq_sps1_seg8_d256_L6 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c6_sps1.json 2>> gpurun_out/c6.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c6_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [(r['kernel'], round(r['avg_launch_ms'],4), round(r['frac'],3)) for r in d['rooflines'] if r['kernel'] in ('gemm_ffn1','gemm_dgrad_ffn2','gemm_qkv','gemm_dgrad_outproj')])
PY
