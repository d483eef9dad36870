# Reproduces the round's one-GPU numbers on a B200 box (≈ 6 minutes): GPU tests, smoke(), the contract bench for Matérn /
# exponential / the reference arm / configs[0] / m = 32, the full-size timings of configs[1], [3], [4], the ncu launch list of
# the bench and one `ncu --set full` capture each of the dominant kernel and of the team-solve instantiation.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_matern_n1e6_m15.json 2>gpurun_out/e1.log; cut -c1-160 gpurun_out/bench_matern_n1e6_m15.json
timeout 300 python bench.py --cov exponential > gpurun_out/bench_exponential_n1e6_m15.json 2>gpurun_out/e2.log; cut -c1-160 gpurun_out/bench_exponential_n1e6_m15.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>gpurun_out/e3.log; cut -c1-160 gpurun_out/bench_reference_arm.json
timeout 300 python bench.py --locations 5000 --cov exponential --steps 1000 --warmup 10 > gpurun_out/bench_exponential_n5000_m15.json 2>gpurun_out/e4.log; cut -c1-160 gpurun_out/bench_exponential_n5000_m15.json
timeout 200 python bench.py --m 32 --locations 500000 --no-cpu-baseline > gpurun_out/bench_matern_n5e5_m32.json 2>gpurun_out/e5.log; cut -c1-160 gpurun_out/bench_matern_n5e5_m32.json
timeout 500 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; tail -6 gpurun_out/bench_configs.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench_matern_n1e6_m15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bf_tri -s 1 -c 1 -o gpurun_out/prof_tri_matern -f python tools/ncu_probe.py 1000000 15 2 3 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bf_tri -s 1 -c 1 -o gpurun_out/prof_tri_m32_exponential -f python tools/ncu_probe.py 500000 32 0 3 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
# summaries: python tools/ncu_summary.py gpurun_out/prof_tri_matern.ncu-rep profiles/<name>.json "<label>"
cat gpurun_out/e*.log | tail -5
