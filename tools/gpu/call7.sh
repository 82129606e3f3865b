mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all7.txt 2>&1; tail -6 gpurun_out/t_all7.txt
python tools/time_gemm.py > gpurun_out/time_gemm_tma7.txt 2>&1; cat gpurun_out/time_gemm_tma7.txt
