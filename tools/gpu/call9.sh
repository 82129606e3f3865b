mkdir -p gpurun_out
python bench.py > gpurun_out/bench9_default.json 2> gpurun_out/bench9_default.err; tail -c 1500 gpurun_out/bench9_default.json; tail -5 gpurun_out/bench9_default.err
