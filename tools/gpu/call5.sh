mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/t_kern5.txt 2>&1; tail -3 gpurun_out/t_kern5.txt
python tools/time_gemm.py > gpurun_out/time_gemm_tma5.txt 2>&1; cat gpurun_out/time_gemm_tma5.txt
python tools/time_qr.py > gpurun_out/time_qr5.txt 2>&1; grep svd gpurun_out/time_qr5.txt
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 24 > gpurun_out/bench5_b24.json 2> gpurun_out/bench5_b24.err; cat gpurun_out/bench5_b24.json
TN_QR_GRAPHS=1 TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 24 > gpurun_out/bench5_b24_graphs.json 2> gpurun_out/bench5_b24_graphs.err; cat gpurun_out/bench5_b24_graphs.json; tail -3 gpurun_out/bench5_b24_graphs.err
TN_GEMM_TMA=0 TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 24 > gpurun_out/bench5_b24_notma.json 2> gpurun_out/bench5_b24_notma.err; cat gpurun_out/bench5_b24_notma.json
