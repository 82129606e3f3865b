mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/t_sharded10.txt 2>&1; tail -4 gpurun_out/t_sharded10.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench10_n2.json 2> gpurun_out/bench10_n2.err; tail -c 2500 gpurun_out/bench10_n2.json; tail -5 gpurun_out/bench10_n2.err
