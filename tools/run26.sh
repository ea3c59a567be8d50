mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r01_final.csv \
    python bench.py > gpurun_out/ncu_ll.log 2>&1
echo "ncu launch list exit $?"
tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_full.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 12 -c 1 -o gpurun_out/sort_pass_r01_final -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
