# multi-GPU bench exactly as the driver launches it; usage: bash tools/job6.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -5 gpurun_out/bench_${N}gpu.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e']); print(json.dumps(d.get('scan'))[:2500])"
