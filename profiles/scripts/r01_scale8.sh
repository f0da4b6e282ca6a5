set -u
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench5_n$N.json 2> gpurun_out/bench5_n$N.err; echo "bench n$N rc=$?"
tail -3 gpurun_out/bench5_n$N.err
cat gpurun_out/bench5_n$N.json | cut -c1-400
