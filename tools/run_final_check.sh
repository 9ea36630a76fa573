python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02g_ref.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02g.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"][:12], d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline_update"]["frac"], d["roofline_update"]["achieved"])
print(d["other_configs"]["C4"]["evals_per_s"], d["other_configs"]["C5"]["ms_per_series"], d["banded_schedule"]["bit_identical_to_dense_schedule"], d["phase_ms"])
print(d["roofline_cholesky"]["frac"], d["roofline_assembly"]["frac"], d["other_configs"]["C3_16_walkers_one_gpu"], d["dense_schedule"]["ms_per_step"], d["roofline_dmma_path"]["frac"], d["gpu_launches"], d["clocks"])
r=json.loads(open("gpurun_out/bench_r02g_ref.json").read().strip().splitlines()[-1]); print(r["value"])
PY
