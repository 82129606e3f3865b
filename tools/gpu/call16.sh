mkdir -p gpurun_out
TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench16_16c.json 2> gpurun_out/bench16_16c.err; cat gpurun_out/bench16_16c.json
TN_BENCH_LITE=1 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench16_4c.json 2> gpurun_out/bench16_4c.err; cat gpurun_out/bench16_4c.json
TN_BENCH_LITE=1 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-extra --batch 16 > gpurun_out/bench16_4c_b16.json 2> gpurun_out/bench16_4c_b16.err; cat gpurun_out/bench16_4c_b16.json
