cd $GRAFT_REPO_ROOT/tools/probes
timeout 200 python attn_tc5_check.py > ../../gpurun_out/c9_check.log 2>&1; echo "check rc=$?"
grep -E "FAIL|shapes ok|FAILED|rror" ../../gpurun_out/c9_check.log | head -20
timeout 120 python attn_tc5_check.py --time-only > ../../gpurun_out/c9_time.log 2>&1; echo "time rc=$?"
AMC_TC5_SPLIT=0 timeout 120 python attn_tc5_check.py --time-only > ../../gpurun_out/c9_time_nosplit.log 2>&1
grep T257 ../../gpurun_out/c9_time.log ../../gpurun_out/c9_time_nosplit.log
