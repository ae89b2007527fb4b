mkdir -p gpurun_out
( lscpu | grep -i -E "numa|^CPU\(s\)|model name"; nvidia-smi topo -m ) > gpurun_out/r02f_topology.txt 2>&1
( time timeout 80 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 3 ) > gpurun_out/r02f_bench_c3_numa_bind.log 2>&1
( time timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 3 --no-numa-bind ) > gpurun_out/r02f_bench_c3_no_bind.log 2>&1
grep -o '"e2e": {[^}]*}' gpurun_out/r02f_bench_c3_numa_bind.log gpurun_out/r02f_bench_c3_no_bind.log
cat gpurun_out/r02f_topology.txt | head -30
