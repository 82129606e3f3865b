mkdir -p gpurun_out
TN_WAIT_MODE=2 TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench15_yield.json 2> gpurun_out/bench15_yield.err; cat gpurun_out/bench15_yield.json
TN_WAIT_MODE=1 TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench15_block.json 2> gpurun_out/bench15_block.err; cat gpurun_out/bench15_block.json
TN_WAIT_MODE=2 TN_BENCH_LITE=1 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench15_yield_4cores.json 2> gpurun_out/bench15_yield_4cores.err; cat gpurun_out/bench15_yield_4cores.json
TN_WAIT_MODE=1 TN_BENCH_LITE=1 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench15_block_4cores.json 2> gpurun_out/bench15_block_4cores.err; cat gpurun_out/bench15_block_4cores.json
python tests/tools/bond_probe.py > gpurun_out/bond_probe.txt 2>&1; tail -12 gpurun_out/bond_probe.txt
