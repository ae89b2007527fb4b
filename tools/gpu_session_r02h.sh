mkdir -p gpurun_out
( time timeout 70 python -m pytest tests/test_load.py tests/test_ingest.py -m gpu -v -x ) > gpurun_out/r02h_ingest_gpu_tests.log 2>&1
echo "pytest_rc=$?" >> gpurun_out/r02h_ingest_gpu_tests.log
grep -E "PASSED|FAILED|ERROR|passed|failed|rc=" gpurun_out/r02h_ingest_gpu_tests.log | tail -40
