mkdir -p gpurun_out
timeout 1500 python tools/bench_cli.py > gpurun_out/r2_bench_cli.log 2> gpurun_out/r2_bench_cli.err; tail -2 gpurun_out/r2_bench_cli.err | cut -c1-300; grep -h '"tool"' gpurun_out/r2_bench_cli.log | tail -1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clock-sampler"
timeout 300 $CMD > gpurun_out/r2_plain_ll.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r2_ncu_ll.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1; head -30 gpurun_out/r02_launches_summary.txt
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clock-sampler"
timeout 300 $CMD > gpurun_out/r2_plain_full.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel|hash_reduce_kernel|expand_scatter_kernel' -c 5 -o gpurun_out/r02_top $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
ncu -i gpurun_out/r02_top.ncu-rep --page details > gpurun_out/r02_top_ncu_details.txt 2>&1
ncu -i gpurun_out/r02_top.ncu-rep --page raw --csv > gpurun_out/r02_top_ncu_raw.csv 2>&1
ls -la gpurun_out/r02_*
