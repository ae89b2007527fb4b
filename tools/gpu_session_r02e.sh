mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02e_box.txt 2>&1
( time timeout 235 python -m pytest tests -m gpu -v -n 6 --durations=15 -rf -p no:cacheprovider ) > gpurun_out/r02e_gpu_tests.log 2>&1
echo "pytest_rc=$?" >> gpurun_out/r02e_gpu_tests.log
( time timeout 40 python __graft_entry__.py --smoke ) > gpurun_out/r02e_smoke.log 2>&1
echo "smoke_rc=$?" >> gpurun_out/r02e_smoke.log
( time timeout 110 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --parity 3 ) > gpurun_out/r02e_bench_c3_n1.log 2>&1
echo "bench_rc=$?" >> gpurun_out/r02e_bench_c3_n1.log
tail -5 gpurun_out/r02e_gpu_tests.log; tail -2 gpurun_out/r02e_smoke.log; tail -c 600 gpurun_out/r02e_bench_c3_n1.log
