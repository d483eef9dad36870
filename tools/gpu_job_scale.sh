# usage: bash tools/gpu_job_scale.sh "8 4"   -- headline bench at the given GPU counts, lines into gpurun_out/r02c_*
mkdir -p gpurun_out
for N in $1; do
  if [ "$N" = 1 ]; then
    timeout 400 python bench.py --steps 20 --warmup 3 2>gpurun_out/scale_${N}_err.log | grep "^{" > gpurun_out/r02c_bench_matern_n1e6_m15_gpus$N.json
  else
    NNGP_TRACE_ITER=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/scale_${N}_err.log | grep "^{" > gpurun_out/r02c_bench_matern_n1e6_m15_gpus$N.json
  fi
  python -c "
import json; d=json.loads(open('gpurun_out/r02c_bench_matern_n1e6_m15_gpus$N.json').read())
print($N, round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'set_data', round(d['e2e']['set_data_s'],4), 'kernel', round(d['roofline']['kernel_ms'],4), d.get('parity') and d['parity']['chain_equal_to_n1'])"
done
