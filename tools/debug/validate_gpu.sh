# One gpurun call that validates a build: GPU tests, smoke(), bench.py, metric counts for tools/make_executed_counts.py,
# one ncu --set full capture of the state kernel and the launch list of bench.py (outputs under gpurun_out/s9_*).
#   gpurun --timeout 2400 -- bash tools/debug/validate_gpu.sh
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s9_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_gputest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s9_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s9_smoke.log
timeout 400 python bench.py > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; echo "bench rc=$?" >> gpurun_out/s9_bench.err
M=smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:pv_state_bits_sorted -s 2 -c 1 --csv --log-file gpurun_out/s9_state_counts.csv python tools/prof_state.py > gpurun_out/s9_ncu1.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:pv_edge -s 8 -c 4 --csv --log-file gpurun_out/s9_edge_counts.csv python tools/prof_edge.py > gpurun_out/s9_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pv_state_bits_sorted -s 2 -c 1 -o gpurun_out/s9_state python tools/prof_state.py > gpurun_out/s9_ncu3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s9_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s9_ncu5.log 2>&1
tail -n 3 gpurun_out/s9_gputest.log; tail -n 2 gpurun_out/s9_smoke.log; cut -c1-200 gpurun_out/s9_bench.json
