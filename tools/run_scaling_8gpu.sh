TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_i8_g8.json 2> gpurun_out/bench_i8_g8.err
$TR bench.py --gpus 8 --steps 20 --warmup 3 --scaling strong > gpurun_out/bench_i8_g8s.json 2> gpurun_out/bench_i8_g8s.err
$TR bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --workload C4 > gpurun_out/bench_i8_c4_g8.json 2> gpurun_out/bench_i8_c4_g8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_i8_g4.json 2> gpurun_out/bench_i8_g4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_i8_g2.json 2> gpurun_out/bench_i8_g2.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i8_g1.json 2> gpurun_out/bench_i8_g1.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload C4 > gpurun_out/bench_i8_c4_g1.json 2> gpurun_out/bench_i8_c4_g1.err
for f in g1 g2 g4 g8 g8s c4_g8 c4_g1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_i8_$f.json").read().strip().splitlines()[-1])
    print("$f", d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), round(d["roofline_cholesky"]["frac"],3))
except Exception as e:
    print("$f", "failed", e)
PY
done
