# 8 x B200 box: weak / strong scaling of the bench workload and C4 (run under `gpurun --gpus 8`)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
TAG=${1:-i8p}
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_g8.json 2> gpurun_out/bench_${TAG}_g8.err
$TR bench.py --gpus 8 --steps 20 --warmup 3 --scaling strong > gpurun_out/bench_${TAG}_g8s.json 2> gpurun_out/bench_${TAG}_g8s.err
$TR bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --workload C4 > gpurun_out/bench_${TAG}_c4_g8.json 2> gpurun_out/bench_${TAG}_c4_g8.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g1.json 2> gpurun_out/bench_${TAG}_g1.err
for f in g1 g8 g8s c4_g8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), round(d["roofline_cholesky"]["frac"],3))
except Exception as e:
    print("$f", "failed", e)
PY
done
