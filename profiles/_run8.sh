mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > gpurun_out/r2/topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/r2/lscpu.txt 2>&1
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2/n8d.json 2> gpurun_out/r2/n8d.err; tail -2 gpurun_out/r2/n8d.err
