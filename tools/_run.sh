python bench.py --steps 5 --no-cpu --no-e2e --no-single | python tools/bsum.py C2finish
( time python bench.py --workload C3 --steps 3 --cpu-queries 2 ) > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo rc=$?; tail -5 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json | python tools/bsum.py C3; cat gpurun_out/bench_c3.json
