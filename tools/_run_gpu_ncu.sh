CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clock-sampler"
timeout 300 $CMD > gpurun_out/plain_full.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel|hash_reduce_kernel' -c 4 -o gpurun_out/r01_final_top $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
