set -u
mkdir -p gpurun_out
python -m pytest tests/test_dist_gpu.py -m gpu -q -rA > gpurun_out/r2_dist_gpu_tests.log 2>&1; echo "dist pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2_dist_gpu_tests.log | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --workload mtan --no-cpu-baseline 2> gpurun_out/b2.err | grep '^{' > gpurun_out/r2_bench_mtan_2gpu.json; echo "bench2 rc=${PIPESTATUS[0]}"
