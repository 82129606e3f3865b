mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/t_all3.txt 2>&1; tail -5 gpurun_out/t_all3.txt
python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench3_fused_b12.json 2> gpurun_out/bench3a.err
TN_QR_APPLY=gemm python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench3_gemm_b12.json 2> gpurun_out/bench3b.err
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 16 > gpurun_out/bench3_fused_b16.json 2> gpurun_out/bench3c.err
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch 24 > gpurun_out/bench3_fused_b24.json 2> gpurun_out/bench3d.err
TN_BENCH_LITE=1 TN_QR_APPLY=gemm python bench.py --steps 2 --warmup 1 --no-extra --batch 24 > gpurun_out/bench3_gemm_b24.json 2> gpurun_out/bench3e.err
python tools/j124_sweep.py 20 8 1 > gpurun_out/j124_sweep_D8.txt 2>&1; tail -3 gpurun_out/j124_sweep_D8.txt
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err
for f in gpurun_out/bench3_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print({k:d.get(k) for k in ['value','ms_per_step','gpu_launches','latency_seconds_single_instance','impl']}, d.get('e2e'))
"; done
