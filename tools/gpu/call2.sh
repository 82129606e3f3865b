mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/t_kern2.txt 2>&1; tail -3 gpurun_out/t_kern2.txt
python tools/time_qr.py > gpurun_out/time_qr2.txt 2>&1
TNAC4O_B200_LIB=$PWD/tools/microbench/lib_phases/libtnac4o_b200.so python tools/microbench/phases.py > gpurun_out/phases2.txt 2>&1
python tools/profile_ops.py > gpurun_out/profile_ops_b.txt 2>&1
python -m pytest tests/test_rmf_gpu.py tests/test_solver_gpu.py -x -q -m gpu > gpurun_out/t_solver2.txt 2>&1; tail -15 gpurun_out/t_solver2.txt
python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 600 gpurun_out/bench2.json
