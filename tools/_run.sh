CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-single"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo rc=$?
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_edge_flat|k_knn_direct|k_csr_finish|k_csr_degree|k_reject_flags' -s 15 -c 6 -o gpurun_out/prof_r1k $CMD > gpurun_out/ncu_full.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_full.log
CMD3="python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-single"
$CMD3 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_edge_flat' -s 3 -c 1 -o gpurun_out/prof_r1k_c3 $CMD3 > gpurun_out/ncu_full3.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_full3.log
