mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_rmf_gpu.py -x -q -m gpu -k "spectrum or config3 or rmf" > gpurun_out/t_18.txt 2>&1; tail -3 gpurun_out/t_18.txt
python bench.py --steps 1 --warmup 1 --no-extra > gpurun_out/bench18.json 2> gpurun_out/bench18.err; python -c "
import json; d=json.load(open('gpurun_out/bench18.json')); r=d['roofline']; print(d['value'], r['achieved'], r['frac'], r['single_instance'], r['peak_all_gpus'])"; tail -3 gpurun_out/bench18.err
