set -x
python -c "import __graft_entry__ as g; g.smoke()" || exit 1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_i8p.json 2> gpurun_out/bench_r01_i8p.err || exit 2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_i8p_ref.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_i8p.log 2> gpurun_out/plain_i8p.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/launches_r01_i8p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i8p_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 40 -c 8 -f -o gpurun_out/prof_r01_i8p_syrk \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i8p_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"i8_slice|trsm_kernel|potrf_diag" -s 60 -c 6 -f -o gpurun_out/prof_r01_i8p_misc \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i8p_misc.log 2>&1
ls -la gpurun_out | tail -8
