for w in 1 2 8 16; do echo "window=$w"; OTTOCOV_SO_NAME=libottocov_w$w.so timeout 300 python tools/bench_sort.py 268435456; done 2>&1
echo "window=4"; timeout 300 python tools/bench_sort.py 268435456
