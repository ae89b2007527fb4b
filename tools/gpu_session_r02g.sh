mkdir -p gpurun_out
( time timeout 95 python -m pytest tests/test_load.py -m gpu -v -x ) > gpurun_out/r02g_load_gpu_tests.log 2>&1
echo "pytest_rc=$?" >> gpurun_out/r02g_load_gpu_tests.log
tail -25 gpurun_out/r02g_load_gpu_tests.log
