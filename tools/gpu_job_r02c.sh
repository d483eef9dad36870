# 8-GPU lease: n = 10^7 headline bench at N = 8 and N = 4
mkdir -p gpurun_out
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --locations 10000000 --steps 10 --warmup 3 2>gpurun_out/bench1e7_${N}_err.log | grep "^{" > gpurun_out/r02_bench_matern_n1e7_m15_gpus$N.json
  cut -c1-220 gpurun_out/r02_bench_matern_n1e7_m15_gpus$N.json; tail -2 gpurun_out/bench1e7_${N}_err.log | cut -c1-300
done
