# Strong (default) and weak scaling of the bench workload over the GPUs of one box.  usage (under gpurun --gpus N):
#   [EXTRA=--no-extras] bash tools/run_scaling_8gpu.sh N [TAG]
N=${1:-8}
TAG=${2:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 20 --warmup 5 ${EXTRA} \
  > gpurun_out/bench_${TAG}_g${N}.json 2> gpurun_out/bench_${TAG}_g${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 ${EXTRA} \
  --scaling weak --no-extras > gpurun_out/bench_${TAG}_g${N}w.json 2>/dev/null
python -c "
import json
for f in ('gpurun_out/bench_${TAG}_g${N}.json', 'gpurun_out/bench_${TAG}_g${N}w.json'):
    d = json.loads(open(f).read()); print(f, d['scaling'], d['value'], d['ms_per_step'])"
