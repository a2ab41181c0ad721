set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/q_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_all.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
