mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bcef.py tests/test_gpu_predict.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/bench_configs.py kfold predict > gpurun_out/bench_configs.log 2>&1; tail -4 gpurun_out/bench_configs.log | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"knn_cb|knn_brute|knn_grid|predict_emit|rep_solve|q_resid_kernel|q_forms_kernel|tri_build|site_reduce" -c 30 -o gpurun_out/prof_others python tools/ncu_others.py > gpurun_out/ncu_others.log 2>&1; tail -2 gpurun_out/ncu_others.log
