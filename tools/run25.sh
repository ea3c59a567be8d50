for b in 64 256 1024; do ./experiments/hash_bench 742000000 $b 1.3 | tail -1; done
./experiments/hash_bench 742000000 256 2.0 | tail -1
