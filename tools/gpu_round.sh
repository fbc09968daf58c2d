#!/usr/bin/env bash
# One gpurun call = one round's GPU evidence (a call costs ~30 s of box time before the command starts, so batch):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round.sh collect r2'     # on the B200 box
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round.sh sanitize r2 memcheck'   # compute-sanitizer subset (one tool per call: memcheck | racecheck)
#   bash tools/gpu_round.sh summarise r2                                                   # back here: -> profiles/
# collect: GPU tests, smoke, the default bench line (+ reference arm), one bench line per other workload, the ncu launch
# list of one training step and one `ncu --set full` capture of that step (each ncu pass only after the same command
# has exited 0 without ncu; nothing printed under ncu is a benchmark number).  Every step has its own timeout so a hang
# cannot eat the box.  Outputs: gpurun_out/<tag>_*.
set -u
MODE=${1:-collect}
TAG=${2:-rX}
OUT=gpurun_out
mkdir -p "$OUT"
WORKLOADS="rawiq_seg16_d128_L6 vit_p4_d128_L6 rawiq_sps1_seg8_d256_L6 rawiq_sps2_seg8_d256_L6 rawiq_seg16_d512_L12 rawiq_conv1d_d128_L6"

# ncu reports with sources are tens of MB each and gpurun copies back at most 64 MiB: keep the raw-metric CSV, drop the report
rep2csv() { [ -s "$1.ncu-rep" ] && ncu -i "$1.ncu-rep" --page raw --csv > "$1.csv" 2>/dev/null; rm -f "$1.ncu-rep"; }

if [ "$MODE" = collect ]; then
  timeout 400 python -m pytest tests -m gpu -q -x > "$OUT/${TAG}_pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 120 python __graft_entry__.py smoke > "$OUT/${TAG}_smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 300 python bench.py > "$OUT/${TAG}_bench_1gpu.json" 2> "$OUT/${TAG}_bench_1gpu.err"; echo "bench rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/${TAG}_bench_reference_arm.json" 2>> "$OUT/${TAG}_bench_1gpu.err"
  echo "reference arm rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 200 python bench.py --dtype fp32 --no-cpu-baseline --steps 10 --warmup 3 > "$OUT/${TAG}_bench_1gpu_fp32.json" 2>> "$OUT/${TAG}_bench_1gpu.err"
  echo "fp32 bench rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  for W in $WORKLOADS; do
    timeout 120 python bench.py --workload "$W" --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_wl_${W}.json" 2>> "$OUT/${TAG}_wl.err"
    echo "workload $W rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  done
  if timeout 120 python tools/one_step.py > "$OUT/${TAG}_one_step_plain.log" 2>&1; then
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file "$OUT/${TAG}_launches_train.csv" python tools/one_step.py > "$OUT/${TAG}_ncu_launch.log" 2>&1
    echo "ncu launch list rc=$?" | tee -a "$OUT/${TAG}_status.txt"
    # the dominant class (weight-gradient GEMM) at the bench batch first -- seconds; then the whole default step
    timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:gemm_tc_kernel -c 8 \
      -o "$OUT/${TAG}_full_wgrad" python tools/one_step.py > "$OUT/${TAG}_ncu_full_wgrad.log" 2>&1
    echo "ncu full wgrad rc=$?" | tee -a "$OUT/${TAG}_status.txt"
    rep2csv "$OUT/${TAG}_full_wgrad"
    # the other classes of the default step (GEMM epilogues, LayerNorm, front end, T = 9 attention): first 70 launches of a
    # step at 8192 frames (the whole step at the bench batch does not finish in 7 minutes under --set full)
    timeout 300 ncu --set full --clock-control none --profile-from-start off -f -c 70 \
      -o "$OUT/${TAG}_full_step" python tools/one_step.py --batch 8192 > "$OUT/${TAG}_ncu_full.log" 2>&1
    echo "ncu full rc=$?" | tee -a "$OUT/${TAG}_status.txt"
    rep2csv "$OUT/${TAG}_full_step"
  else
    echo "one_step.py failed without ncu: no ncu passes" | tee -a "$OUT/${TAG}_status.txt"
  fi
  # the tcgen05 attention kernels on the raw-IQ SPS-1 / SPS-2 shapes (T = 129 / 257): launch list of the step + full capture
  # of the attention kernels only
  for W in rawiq_sps1_seg8_d256_L6 rawiq_sps2_seg8_d256_L6; do
    if timeout 120 python tools/one_step.py --workload $W > "$OUT/${TAG}_one_step_${W}.log" 2>&1; then
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file "$OUT/${TAG}_launches_${W}.csv" python tools/one_step.py --workload $W > "$OUT/${TAG}_ncu_launch_${W}.log" 2>&1
      echo "ncu launch list $W rc=$?" | tee -a "$OUT/${TAG}_status.txt"
      timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:attn_ \
        -o "$OUT/${TAG}_full_attn_${W}" python tools/one_step.py --workload $W > "$OUT/${TAG}_ncu_full_${W}.log" 2>&1
      echo "ncu full attention $W rc=$?" | tee -a "$OUT/${TAG}_status.txt"
      rep2csv "$OUT/${TAG}_full_attn_${W}"
    fi
  done
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/${TAG}_nvidia_smi.csv" 2>&1
  cat "$OUT/${TAG}_status.txt"
elif [ "$MODE" = refresh ]; then
  # after a kernel change late in a round: tests, smoke, the bench lines, the default launch list and a full capture of the
  # row-epilogue GEMMs only (the other captures of `collect` stay valid)
  timeout 400 python -m pytest tests -m gpu -q -x > "$OUT/${TAG}_pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 120 python __graft_entry__.py smoke > "$OUT/${TAG}_smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 300 python bench.py > "$OUT/${TAG}_bench_1gpu.json" 2> "$OUT/${TAG}_bench_1gpu.err"; echo "bench rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  for W in $WORKLOADS; do
    timeout 120 python bench.py --workload "$W" --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_wl_${W}.json" 2>> "$OUT/${TAG}_wl.err"
    echo "workload $W rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  done
  for B in 256 2048; do
    timeout 120 python bench.py --workload rawiq_seg16_d128_L6 --batch $B --graph --steps 20 --warmup 5 --no-cpu-baseline \
      > "$OUT/${TAG}_wl_rawiq_seg16_d128_L6_B${B}_graph.json" 2>> "$OUT/${TAG}_wl.err"
    echo "graph step B=$B rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  done
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file "$OUT/${TAG}_launches_train.csv" python tools/one_step.py > "$OUT/${TAG}_ncu_launch.log" 2>&1
  echo "ncu launch list rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  timeout 240 ncu --set full --clock-control none --profile-from-start off -f -k regex:gemm_tc_row_kernel -c 42 \
    -o "$OUT/${TAG}_full_rowgemm" python tools/one_step.py --batch 8192 > "$OUT/${TAG}_ncu_full_rowgemm.log" 2>&1
  echo "ncu full row GEMMs rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  rep2csv "$OUT/${TAG}_full_rowgemm"
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/${TAG}_nvidia_smi.csv" 2>&1
  cat "$OUT/${TAG}_status.txt"
elif [ "$MODE" = sanitize ]; then
  # memcheck + racecheck over the op-level tests at their smallest shapes (the sanitizer slows kernels
  # 10-100x: a bounded subset, each tool under its own timeout).  gpurun --timeout 900 -- 'bash tools/gpu_round.sh sanitize r2'
  SEL="test_attention_fwd_bwd and (3-9-8-32 or 2-65-8-16 or 1-257-8-32) or test_long_attention_fwd_bwd and 1-300-2-16 or test_cls_row and (3-9-8-32 or 2-289-2-16) or test_layernorm_fwd_bwd and 64-16 or test_gemm_epilogues"
  # one tool per gpurun call (B200_PROFILING.md: several sanitizer tools in one call have left the GPU unusable)
  for TOOL in ${3:-memcheck}; do
    timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file "$OUT/${TAG}_sanitizer_${TOOL}.log" \
      python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL" > "$OUT/${TAG}_sanitizer_${TOOL}_pytest.log" 2>&1
    echo "compute-sanitizer $TOOL rc=$?" | tee -a "$OUT/${TAG}_status.txt"
    tail -5 "$OUT/${TAG}_sanitizer_${TOOL}.log"
  done
elif [ "$MODE" = scale ]; then
  # multi-GPU evidence on one box: `gpurun --gpus N -- bash tools/gpu_round.sh scale r2 "1 2"` (N = the largest count listed).
  # Per count: the raw-IQ SPS-2 and cfg-1 workloads at 1024 and 256 frames per GPU (BASELINE configs[2] / [0]), a strong-
  # scaling line of the default workload (32768 frames in total), the DP parity check, and -- at 8 GPUs -- the tuner grid.
  NS=${3:-"1 2"}
  run_n() {   # N, out file, bench args...
    local n=$1 out=$2; shift 2
    if [ "$n" = 1 ]; then timeout 200 python bench.py --gpus 1 "$@" > "$out" 2>> "$OUT/${TAG}_scale.err"
    else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
           --master-port $((29500 + n)) bench.py --gpus "$n" "$@" > "$out" 2>> "$OUT/${TAG}_scale.err"; fi
    echo "scale n=$n $* rc=$?" | tee -a "$OUT/${TAG}_status.txt"
  }
  for N in $NS; do
    for W in rawiq_sps2_seg8_d256_L6 rawiq_seg16_d128_L6; do
      for B in 1024 256; do
        run_n "$N" "$OUT/${TAG}_scale_${W}_B${B}_${N}gpu.json" --workload "$W" --batch "$B" --steps 20 --warmup 5 --no-cpu-baseline
      done
    done
    run_n "$N" "$OUT/${TAG}_strong_vit_p16_G32768_${N}gpu.json" --global-batch 32768 --steps 20 --warmup 5 --no-cpu-baseline
    if [ "$N" -gt 1 ]; then
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + N)) \
        tools/dp_check.py > "$OUT/${TAG}_dp_check_${N}gpu.log" 2>&1; echo "dp_check n=$N rc=$?" | tee -a "$OUT/${TAG}_status.txt"
      tail -2 "$OUT/${TAG}_dp_check_${N}gpu.log"
    fi
    if [ "$N" = 8 ]; then
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
        -m vit_vs_raw_iq_b200.tuning --particles 16 --iters 2 > "$OUT/${TAG}_tuning_8gpu.log" 2>&1; echo "tuning x8 rc=$?" | tee -a "$OUT/${TAG}_status.txt"
      tail -3 "$OUT/${TAG}_tuning_8gpu.log"
    fi
  done
  cat "$OUT/${TAG}_status.txt"
elif [ "$MODE" = summarise ]; then
  cp "$OUT/${TAG}_bench_1gpu.json" "profiles/${TAG}_bench_1gpu.json"
  cp "$OUT/${TAG}_bench_reference_arm.json" "profiles/${TAG}_bench_reference_arm.json"
  for W in $WORKLOADS; do
    [ -s "$OUT/${TAG}_wl_${W}.json" ] && cp "$OUT/${TAG}_wl_${W}.json" "profiles/workloads/${TAG}_wl_${W}.json"
  done
  [ -s "$OUT/${TAG}_launches_train.csv" ] && cp "$OUT/${TAG}_launches_train.csv" "profiles/${TAG}_launches_train.csv" && \
    python tools/ncu_summary.py launches "$OUT/${TAG}_launches_train.csv" "profiles/${TAG}_launches_train.md" \
      --title "${TAG}: ncu launch list of one training step (tools/one_step.py, default workload)"
  cp "$OUT/${TAG}_bench_1gpu_fp32.json" "profiles/${TAG}_bench_1gpu_fp32.json"
  [ -s "$OUT/${TAG}_full_step.csv" ] && \
    python tools/ncu_summary.py full "$OUT/${TAG}_full_step.csv" "profiles/${TAG}_ncu_full_step.json" --batch 8192
  # (after the whole-step summary: the bench-batch capture of the dominant class is the one bench.py reads)
  [ -s "$OUT/${TAG}_full_wgrad.csv" ] && \
    python tools/ncu_summary.py full "$OUT/${TAG}_full_wgrad.csv" "profiles/${TAG}_ncu_full_wgrad.json"
  for W in rawiq_sps1_seg8_d256_L6 rawiq_sps2_seg8_d256_L6; do
    [ -s "$OUT/${TAG}_launches_${W}.csv" ] && python tools/ncu_summary.py launches "$OUT/${TAG}_launches_${W}.csv" \
      "profiles/${TAG}_launches_train_${W}.md" --title "${TAG}: ncu launch list of one training step (tools/one_step.py --workload $W)"
    [ -s "$OUT/${TAG}_full_attn_${W}.csv" ] && \
      python tools/ncu_summary.py full "$OUT/${TAG}_full_attn_${W}.csv" "profiles/${TAG}_ncu_full_attn_${W}.json" --workload $W
  done
  tail -3 "$OUT/${TAG}_pytest_gpu.log"
else
  echo "usage: $0 collect|sanitize|summarise <tag>"; exit 2
fi
