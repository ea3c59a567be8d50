# round 2, final N=1 evidence run of the last build: GPU tests, smoke, default bench, launch list + ncu of the top kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2g_pytest.log 2>&1; tail -4 gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_cooc_n1.log 2> gpurun_out/r2g_bench_cooc_n1.err; echo "bench exit $?"; tail -2 gpurun_out/r2g_bench_cooc_n1.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2g_bench_cooc_n1.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clock-sampler"
timeout 300 $CMD > gpurun_out/r2g_plain_ll.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r2g_ncu_ll.log 2>&1
python tools/summarize_launches.py gpurun_out/r02b_launches.csv > gpurun_out/r02b_launches_summary.txt 2>&1; head -24 gpurun_out/r02b_launches_summary.txt
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clock-sampler --no-streamed-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel|hash_reduce_buckets_kernel|expand_scatter_kernel' -c 4 -o gpurun_out/r02b_top $CMD > gpurun_out/r2g_ncu_full.log 2>&1
tail -2 gpurun_out/r2g_ncu_full.log
ncu -i gpurun_out/r02b_top.ncu-rep --page details > gpurun_out/r02b_top_ncu_details.txt 2>&1
ncu -i gpurun_out/r02b_top.ncu-rep --page raw --csv > gpurun_out/r02b_top_ncu_raw.csv 2>&1
