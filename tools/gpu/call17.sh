mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all17.txt 2>&1; tail -3 gpurun_out/t_all17.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench17_default.json 2> gpurun_out/bench17_default.err; tail -c 300 gpurun_out/bench17_default.json; tail -3 gpurun_out/bench17_default.err
TN_BENCH_LITE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench17.csv python bench.py --steps 1 --warmup 0 --batch 1 --no-extra > gpurun_out/ncu_bench17.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_bench17.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 3000: TN_BENCH_LITE=1 python bench.py --steps 1 --warmup 0 --batch 1 --no-extra (first 3000 launches of one instance, round 2 final build)" > gpurun_out/launches_bench17_summary.csv; head -30 gpurun_out/launches_bench17_summary.csv
