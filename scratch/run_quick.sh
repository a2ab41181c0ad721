set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_tests.log
python tools/site_bench.py --only "${1:-}" > gpurun_out/q_site.txt 2>&1; echo "site rc=$?"
python bench.py --workload mtan --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_mtan.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
