mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_dropin.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dropin.log
tail -25 gpurun_out/pytest_dropin.log
