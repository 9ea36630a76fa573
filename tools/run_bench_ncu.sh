# Round-2 evidence run (one B200, under gpurun): tests, bench (both arms), ncu launch list, full-set captures.
# The two full-set reports must stay below 64 MiB together (gpurun does not copy a larger gpurun_out/ back).
# Every ncu command runs only after the same command has exited 0 without the profiler.
set -x
TAG=${TAG:-r02}
python -c "import __graft_entry__ as g; g.smoke()" || exit 1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_${TAG}.log 2> gpurun_out/plain_${TAG}.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 40 -c 2 -f -o gpurun_out/prof_${TAG}_syrk \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"trsm_kernel|potrf_diag|assemble_fast" -s 92 -c 3 -f -o gpurun_out/prof_${TAG}_misc \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${TAG}_misc.log 2>&1
for r in syrk misc; do ncu -i gpurun_out/prof_${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$r.raw.csv 2>/dev/null; done
# keep the copy-back below the limit: the raw pages survive even if a report has to go
if [ $(du -sm gpurun_out | cut -f1) -ge 62 ]; then rm -f gpurun_out/prof_${TAG}_misc.ncu-rep; fi
du -sh gpurun_out; ls -la gpurun_out | tail -12
