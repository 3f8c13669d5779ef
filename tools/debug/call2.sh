cd $GRAFT_REPO_ROOT
timeout 300 python tools/probes/edge_cert_ab.py > gpurun_out/s4_cert2_ab.log 2>&1; echo "rc=$?" >> gpurun_out/s4_cert2_ab.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "edge or motion or cert" > gpurun_out/s4_cert2_test.log 2>&1; echo "rc=$?" >> gpurun_out/s4_cert2_test.log
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,sm__icc_request_hit_rate.pct
timeout 300 ncu --metrics $M --clock-control none -k regex:pv_edge -s 8 -c 4 --csv --log-file gpurun_out/s4_cert2_ncu.csv python tools/prof_edge.py > gpurun_out/s4_cert2_ncu.log 2>&1
cat gpurun_out/s4_cert2_ab.log; tail -n 4 gpurun_out/s4_cert2_test.log
