mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench14_n8.json 2> gpurun_out/bench14_n8.err; tail -c 600 gpurun_out/bench14_n8.json; tail -5 gpurun_out/bench14_n8.err
