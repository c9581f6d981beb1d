# final: tests + bench (+ reference arm) + launch list.   gpurun --timeout 1500 -- 'bash tools/r2_call18.sh'
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_gputests.log
timeout 400 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python profiles/pick.py < gpurun_out/r2i_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --min-warmup-s 0 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/r2i_bench_ncu.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
name=r02_ncu_persist_c4
AB_REPS=2 AB_CONFIG=c4 FDB_VERBOSE=1 timeout 300 $NCU -k regex:k_fused_persist -s 4 -c 1 -o gpurun_out/$name python tools/ab_assembly.py > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep; echo "ncu done"
