# 2-GPU evidence: NCCL parity tests, then the BASELINE configs[4] data-parallel batch sweep (per-GPU batch = global / 2)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_dist_gpu.py -m gpu -q > gpurun_out/r2_dist_gpu_tests.log 2>&1; echo "dist pytest rc=$?"; tail -3 gpurun_out/r2_dist_gpu_tests.log
: > gpurun_out/r2_batch_sweep_2gpu.jsonl
run() { # workload, per-GPU batch
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus 2 --steps 5 --warmup 3 --workload "$1" --batch "$2" --no-cpu-baseline 2> gpurun_out/sweep_err.log | grep '^{' >> gpurun_out/r2_batch_sweep_2gpu.jsonl
  echo "sweep $1 B=$2 rc=${PIPESTATUS[0]}"
}
run mtan 32
run mtan 64
run mtan 128
run csnet 32
run csnet 64
run csnet 128
run csnet 256
wc -l gpurun_out/r2_batch_sweep_2gpu.jsonl
