# Round-2 evidence on one B200: GPU tests, per-site bench, the bench line, the ncu launch list of the site bench and one
# --set full capture of the dominant kernels.  (The ncu launch list of a whole eager bench step --
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv python bench.py --workload mtan --steps 2 \
#       --warmup 3 --no-cpu-baseline --no-graph
# -- takes ~15 GPU-minutes for its ~14 000 launches; profiles/r2_ncu_launches_bench_mtan_summary.txt is its digest.)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_gpu_tests.log
python tools/site_bench.py --json gpurun_out/r2_site_bench.json > gpurun_out/r2_site_bench.txt 2>&1; echo "site rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_all.json 2> gpurun_out/r2_bench_all.err; echo "bench rc=$?"
python scratch/step_breakdown.py --model mtan 2>/dev/null | cut -c1-250 > gpurun_out/r2_step_breakdown_mtan.txt
python scratch/step_breakdown.py --model csnet 2>/dev/null | cut -c1-250 > gpurun_out/r2_step_breakdown_csnet.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_ncu_launches_site_bench.csv python tools/site_bench.py --once > gpurun_out/ncu1.log 2>&1; echo "ncu site launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gate_tc_|head_ce_tc|xstitch_|bn_|up2_" -c 40 -o gpurun_out/r2_full python tools/site_bench.py --once --profile-set > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
