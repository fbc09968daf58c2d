cd $GRAFT_REPO_ROOT/tools/probes
timeout 200 python attn_tc5_check.py --bwd > ../../gpurun_out/c2_check.log 2>&1; echo "check rc=$?"
grep -E "FAIL|shapes ok|FAILED|error" ../../gpurun_out/c2_check.log | head -20
timeout 120 python attn_tc5_check.py --bwd --time-only > ../../gpurun_out/c2_time.log 2>&1; echo "time rc=$?"
cat ../../gpurun_out/c2_time.log | tail -8
for sh in "1024 129 8 32" "1024 128 8 32" "512 257 8 32"; do
  AMC_TC5_TRACE=1 timeout 60 python attn_tc5_trace.py $sh --bwd > ../../gpurun_out/c2_trace_$(echo $sh | tr ' ' '_').log 2>&1
done
