set -x
M=smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
python tools/prof_state.py > gpurun_out/plain_state.log 2>&1 && ncu --metrics $M --clock-control none -k regex:pv_state_bits_sorted -s 2 -c 1 --csv --log-file gpurun_out/${TAG:-r2m}_state_counts.csv python tools/prof_state.py > /dev/null 2>&1
python tools/prof_edge.py > gpurun_out/plain_edge.log 2>&1 && ncu --metrics $M --clock-control none -k regex:pv_edge_kernel -s 1 -c 1 --csv --log-file gpurun_out/${TAG:-r2m}_edge_counts.csv python tools/prof_edge.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pv_state_bits_sorted -s 2 -c 1 -o gpurun_out/${TAG:-r2m}_state -f python tools/prof_state.py > gpurun_out/ncu_state.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pv_edge_kernel -s 1 -c 1 -o gpurun_out/${TAG:-r2m}_edge -f python tools/prof_edge.py > gpurun_out/ncu_edge.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG:-r2m}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -20
