# Round-2 evidence run (one B200, under gpurun): tests, bench (both arms), ncu launch list, full-set captures.
# Every ncu command runs only after the same command has exited 0 without the profiler.
set -x
TAG=${TAG:-r02}
python -c "import __graft_entry__ as g; g.smoke()" || exit 1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_${TAG}.log 2> gpurun_out/plain_${TAG}.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 40 -c 6 -f -o gpurun_out/prof_${TAG}_syrk \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"trsm_kernel|potrf_diag|assemble_fast" -s 30 -c 6 -f -o gpurun_out/prof_${TAG}_misc \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_misc.log 2>&1
ls -la gpurun_out | tail -8
