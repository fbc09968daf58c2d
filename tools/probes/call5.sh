cd $GRAFT_REPO_ROOT
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c5_$tag.json 2>> gpurun_out/c5.err; echo "$tag rc=$?"
}
run base AMC_X=1
run s70 AMC_MMA_SMEM_KB=70
run s36 AMC_MMA_SMEM_KB=36 AMC_MMA_FMIN=1
run s70b AMC_MMA_SMEM_KB=100 AMC_MMA_SMEM_KB_BWD=70
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [(r['kernel'], round(r['avg_launch_ms'],4), round(r['frac'],3)) for r in d['rooflines'] if r['kernel'] in ('gemm_ffn1','gemm_qkv','attn_bwd','attn_fwd')])
PY
