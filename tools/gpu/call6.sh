mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "qr" > gpurun_out/t_qr6.txt 2>&1; tail -3 gpurun_out/t_qr6.txt
python tools/time_qr.py > gpurun_out/time_qr6_lean.txt 2>&1; grep qr gpurun_out/time_qr6_lean.txt
TN_QR_PANEL=old python tools/time_qr.py > gpurun_out/time_qr6_old.txt 2>&1; grep qr gpurun_out/time_qr6_old.txt
python tools/profile_ops.py > gpurun_out/profile_ops_c.txt 2>&1; head -8 gpurun_out/profile_ops_c.txt; tail -2 gpurun_out/profile_ops_c.txt
python tests/tools/parity_probe.py > gpurun_out/parity_probe.txt 2>&1; cat gpurun_out/parity_probe.txt
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 24 > gpurun_out/bench6_b24.json 2> gpurun_out/bench6_b24.err; cat gpurun_out/bench6_b24.json
