mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_sharded_gpu.py tests/test_rmf_gpu.py -x -q -m gpu > gpurun_out/t_solver8.txt 2>&1; tail -4 gpurun_out/t_solver8.txt
python tools/probe_configs.py > gpurun_out/probe_configs8.txt 2>&1; cat gpurun_out/probe_configs8.txt
TN_MARGINALS=cta python tools/probe_configs.py > gpurun_out/probe_configs8_cta.txt 2>&1; cat gpurun_out/probe_configs8_cta.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_gibbs_row8.csv python tools/run_one.py gibbs > gpurun_out/ncu_gibbs8.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_gibbs_row8.csv "gibbs row 1e5 samples r2" | head -20
