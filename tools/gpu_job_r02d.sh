mkdir -p gpurun_out
for cov in 2 0; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:bf_tri -s 1 -c 1 -o gpurun_out/prof_tri_final_cov$cov python tools/ncu_probe.py 1000000 15 $cov 3 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
done
for m in 20 24 26; do for t in 1 0; do echo "m=$m tri=$t"; NNGP_TRI=$t timeout 100 python tools/quick_bench.py 300000 $m 2>&1 | grep -E "batch=1"; done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_matern_n1e6_m15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; tail -1 gpurun_out/b_ncu.log | cut -c1-100
