python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
for PC in 18 22; do python bench.py --steps 5 --knn-per-cell $PC --no-cpu --no-e2e --no-single | python tools/bsum.py direct3_pc$PC; done
CMD="python bench.py --steps 2 --warmup 3 --queries 512 --no-cpu --no-e2e --no-single"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:k_knn -c 80 --csv --log-file gpurun_out/launches_r1j.csv $CMD > gpurun_out/ncu_list.log 2>&1
