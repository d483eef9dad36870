# Round-2 closing measurements on one B200 (bench lines, launch list, ncu captures); everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r02e_bench_matern_n1e6_m15_gpus1.json 2>gpurun_out/e1.log; cut -c1-160 gpurun_out/r02e_bench_matern_n1e6_m15_gpus1.json
timeout 300 python bench.py --cov exponential > gpurun_out/r02e_bench_exponential_n1e6_m15.json 2>gpurun_out/e2.log; cut -c1-160 gpurun_out/r02e_bench_exponential_n1e6_m15.json
timeout 300 python bench.py --locations 5000 --cov exponential --steps 1000 --warmup 10 > gpurun_out/r02e_bench_exponential_n5000_m15.json 2>gpurun_out/e4.log; cut -c1-160 gpurun_out/r02e_bench_exponential_n5000_m15.json
timeout 200 python bench.py --m 32 --locations 500000 --no-cpu-baseline > gpurun_out/r02e_bench_matern_n5e5_m32.json 2>gpurun_out/e5.log; cut -c1-160 gpurun_out/r02e_bench_matern_n5e5_m32.json
timeout 500 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; tail -4 gpurun_out/bench_configs.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02e_launches_bench_matern_n1e6_m15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; tail -1 gpurun_out/b_ncu.log | cut -c1-100
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bf_tri -s 1 -c 1 -o gpurun_out/prof_tri_r02e_matern -f python tools/ncu_probe.py 1000000 15 2 3 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bf_tri -s 1 -c 1 -o gpurun_out/prof_tri_r02e_m32_exponential -f python tools/ncu_probe.py 500000 32 0 3 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
cat gpurun_out/e*.log | tail -5
