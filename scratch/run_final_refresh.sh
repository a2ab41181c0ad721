set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_gpu_tests.log
python tools/site_bench.py --json gpurun_out/r2_site_bench.json > gpurun_out/r2_site_bench.txt 2>&1; echo "site rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_all.json 2> gpurun_out/r2_bench_all.err; echo "bench rc=$?"
