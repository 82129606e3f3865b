mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "svd" > gpurun_out/t_svd11.txt 2>&1; tail -2 gpurun_out/t_svd11.txt
TN_THROUGHPUT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_solver_gpu.py -x -q -m gpu -k "svd or config2 or config4_L2048_M1024" > gpurun_out/t_svd11b.txt 2>&1; tail -2 gpurun_out/t_svd11b.txt
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench11_tp1.json 2> gpurun_out/bench11_tp1.err; cat gpurun_out/bench11_tp1.json
TN_THROUGHPUT=0 TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench11_tp0.json 2> gpurun_out/bench11_tp0.err; cat gpurun_out/bench11_tp0.json
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 32 > gpurun_out/bench11_tp1_b32.json 2> gpurun_out/bench11_tp1_b32.err; cat gpurun_out/bench11_tp1_b32.json
python tools/instances10.py 32 0 > gpurun_out/instances10_D32.json 2> gpurun_out/instances10_D32.err; cat gpurun_out/instances10_D32.err
