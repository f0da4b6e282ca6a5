set -u
mkdir -p gpurun_out
for cfg in 4 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$cfg bench.py --gpus 2 --config $cfg --steps 3 --warmup 3 > gpurun_out/bench14_cfg${cfg}_n2.json 2> gpurun_out/bench14_cfg${cfg}_n2.err; echo "cfg$cfg n2 rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/bench14_cfg${cfg}_n2.json').read().strip().splitlines()[-1]); print($cfg, d['value'], d['ms_per_step'], d['e2e']['value'], d['davidson'])
P
done
