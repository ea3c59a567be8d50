set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clock-sampler"
timeout 300 $CMD > gpurun_out/plain_full.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_onepass|hash_reduce|expand_kernel|rs_onesweep' -s 15 -c 7 -o gpurun_out/r01_hash_path $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
