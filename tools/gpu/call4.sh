mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/t_gemm4.txt 2>&1; tail -3 gpurun_out/t_gemm4.txt
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_rmf_gpu.py tests/test_zz_encodings_gpu.py -x -q -m gpu > gpurun_out/t_solver4.txt 2>&1; tail -12 gpurun_out/t_solver4.txt
python tools/run_one.py gemm > gpurun_out/run_gemm4.txt 2>&1
TN_GEMM_TMA=0 python tools/time_gemm.py > gpurun_out/time_gemm_cpasync.txt 2>&1
python tools/time_gemm.py > gpurun_out/time_gemm_tma.txt 2>&1
cat gpurun_out/time_gemm_cpasync.txt gpurun_out/time_gemm_tma.txt
for B in 24 32 48; do TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch $B > gpurun_out/bench4_b$B.json 2> gpurun_out/bench4_b$B.err; cat gpurun_out/bench4_b$B.json; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -c 2 -f -o gpurun_out/r2_gemm_tma python tools/run_one.py gemm > gpurun_out/ncu_gemm_tma.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_panel_reg_kernel -c 2 -f -o gpurun_out/r2_qr_panel python tools/run_one.py qr > gpurun_out/ncu_qr_panel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_cluster_kernel -c 1 -f -o gpurun_out/r2_jacobi python tools/run_one.py svd > gpurun_out/ncu_jacobi.log 2>&1
ls -la gpurun_out/*.ncu-rep
