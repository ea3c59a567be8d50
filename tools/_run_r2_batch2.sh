mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 --durations=25 -k "fused or scatter or asymmetric or hash_reduce or exchange_first or reduce_pairs or sort_ or golden or random_vs_oracle" > gpurun_out/r2b2_pytest.log 2>&1; tail -45 gpurun_out/r2b2_pytest.log
