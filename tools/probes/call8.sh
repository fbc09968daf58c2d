cd $GRAFT_REPO_ROOT
for K in 256 4096; do
for W in vit_p16_d256_L6 rawiq_seg16_d128_L6; do
AMC_LN8_MAXK=$K timeout 120 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c8_${K}_$W.json 2>> gpurun_out/c8.err; echo "rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c8_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [(r['kernel'], round(r['avg_launch_ms'],4), round(r['frac'],3)) for r in d['rooflines'] if r['kernel'] in ('gemm_outproj_ln','gemm_ffn2_ln')])
PY
