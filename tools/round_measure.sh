mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r02_bench_matern_n1e6_m15.json 2>gpurun_out/e1.log; cut -c1-120 gpurun_out/r02_bench_matern_n1e6_m15.json
timeout 300 python bench.py --cov exponential > gpurun_out/r02_bench_exponential_n1e6_m15.json 2>gpurun_out/e2.log; cut -c1-120 gpurun_out/r02_bench_exponential_n1e6_m15.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>gpurun_out/e3.log; cut -c1-120 gpurun_out/r02_bench_reference_arm.json
timeout 300 python bench.py --locations 5000 --cov exponential --steps 1000 --warmup 10 > gpurun_out/r02_bench_exponential_n5000_m15.json 2>gpurun_out/e4.log; cut -c1-120 gpurun_out/r02_bench_exponential_n5000_m15.json
timeout 300 python bench.py --locations 10000000 --no-cpu-baseline > gpurun_out/r02_bench_matern_n1e7_m15.json 2>gpurun_out/e5.log; cut -c1-120 gpurun_out/r02_bench_matern_n1e7_m15.json
timeout 500 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; tail -8 gpurun_out/bench_configs.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_matern_n1e6_m15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bf_ws -s 1 -c 1 -o gpurun_out/prof_ws_matern_final python tools/ncu_probe.py 1000000 15 2 3 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
cat gpurun_out/e*.log | tail -5
