mkdir -p gpurun_out
./tools/microbench/cluster_sync > gpurun_out/cluster_sync.txt 2>&1
TNAC4O_B200_LIB=$PWD/tools/microbench/lib_phases/libtnac4o_b200.so python tools/microbench/phases.py > gpurun_out/phases_new.txt 2>&1
TN_SVD_V1=1 TNAC4O_B200_LIB=$PWD/tools/microbench/lib_phases/libtnac4o_b200.so python tools/microbench/phases.py > gpurun_out/phases_v1.txt 2>&1
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "svd or qr" > gpurun_out/t_kern.txt 2>&1
python tools/profile_ops.py > gpurun_out/profile_ops_a.txt 2>&1
tail -3 gpurun_out/t_kern.txt
