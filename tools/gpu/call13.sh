mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all13.txt 2>&1; tail -4 gpurun_out/t_all13.txt
python __graft_entry__.py smoke > gpurun_out/smoke13.txt 2>&1; tail -2 gpurun_out/smoke13.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench13_default.json 2> gpurun_out/bench13_default.err; tail -c 400 gpurun_out/bench13_default.json; tail -3 gpurun_out/bench13_default.err
