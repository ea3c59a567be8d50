mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.log; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_ref.log
tail -2 gpurun_out/bench_ref.log
# launch list (full size, short run)
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_ll.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
echo "ncu launch list exit $?"
# full capture of the dominant kernel at 2M sessions
timeout 600 python bench.py --sessions 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_full.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep -s 6 -c 2 -o gpurun_out/sort_pass_r01 -f \
    python bench.py --sessions 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
