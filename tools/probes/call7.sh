cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for W in vit_p16_d256_L6 rawiq_sps1_seg8_d256_L6 rawiq_seg16_d128_L6 rawiq_seg16_d512_L12; do
timeout 120 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c7_$W.json 2>> gpurun_out/c7.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [(r['kernel'], round(r['avg_launch_ms'],4), round(r['frac'],3)) for r in d['rooflines'] if r['kernel'] in ('gemm_outproj_ln','gemm_ffn2_ln','gemm_dgrad_ffn2','gemm_ffn1')])
PY
