mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_solver_gpu.py -x -q -m gpu -k "ten_instances" > gpurun_out/t_ten12.txt 2>&1; tail -3 gpurun_out/t_ten12.txt
run() { name=$1; shift; env "$@" TN_BENCH_LITE=1 python bench.py --steps 2 --warmup 1 --no-extra --batch ${B:-32} > gpurun_out/bench12_$name.json 2> gpurun_out/bench12_$name.err; echo $name; cat gpurun_out/bench12_$name.json; }
run base X=1
run wy74 TN_WY_CTAS=74
run wy37 TN_WY_CTAS=37
run sk37 TN_GEMM_SPLITK_CTAS=37
run sk37wy74 TN_GEMM_SPLITK_CTAS=37 TN_WY_CTAS=74
B=40 run b40 X=1
