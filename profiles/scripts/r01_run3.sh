set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi3.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/bench3_n2.json 2> gpurun_out/bench3_n2.err; echo "bench n2 rc=$?"
python bench.py > gpurun_out/bench3_n1.json 2> gpurun_out/bench3_n1.err; echo "bench n1 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err; echo "ref rc=$?"
