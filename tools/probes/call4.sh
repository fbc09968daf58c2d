cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for W in vit_p16_d256_L6 vit_p4_d128_L6 rawiq_seg16_d512_L12 rawiq_sps1_seg8_d256_L6; do
  timeout 120 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c4_wl_$W.json 2>> gpurun_out/c4.err; echo "$W rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c4_wl_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), [(r['kernel'], round(r['avg_launch_ms'],4), round(r['frac'],3)) for r in d['rooflines'] if r['kernel'] in ('gemm_ffn1','gemm_qkv','gemm_dgrad_outproj','attn_bwd','attn_fwd')])
PY
