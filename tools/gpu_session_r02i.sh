mkdir -p gpurun_out
( time timeout 28 python -m pytest tests/test_coarsen_gpu.py tests/test_load.py -m gpu -q -x -k "general_aggregation or general_merge or extra_columns or unordered or pairs_fields" ) > gpurun_out/r02i_aggregate_gpu_tests.log 2>&1
echo "pytest_rc=$?" >> gpurun_out/r02i_aggregate_gpu_tests.log
tail -6 gpurun_out/r02i_aggregate_gpu_tests.log
