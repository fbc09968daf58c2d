cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 120 python bench.py --workload rawiq_seg16_d512_L12 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c11_d512.json 2>> gpurun_out/c11.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c11_d512.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], [(r['kernel'], r['launches_per_step'], round(r['avg_launch_ms'],4), round(r['frac'],3)) for r in d['rooflines'][:12]])
PY
