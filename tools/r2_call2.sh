# Round-2 check of the row-wise pattern build + the P1-tetrahedra launch bounds.   gpurun --timeout 1200 -- 'bash tools/r2_call2.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rowwise or bit_identical or operators or reproducible" > gpurun_out/r2b_quick.log 2>&1; echo "quick rc=$?"; tail -5 gpurun_out/r2b_quick.log
timeout 400 python bench.py --no-extra > gpurun_out/r2b_bench_rows.json 2> gpurun_out/r2b_bench_rows.err; echo "bench rows rc=$?"
FDB_PATTERN_SORT=1 timeout 400 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_sort.json 2> gpurun_out/r2b_bench_sort.err; echo "bench sort rc=$?"
python profiles/pick.py < gpurun_out/r2b_bench_rows.json; python profiles/pick.py < gpurun_out/r2b_bench_sort.json
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_gputests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --min-warmup-s 0 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/r2b_bench_ncu.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
prof() { name=$1; kre=$2; skip=$3; cnt=$4; shift 4; env "$@" AB_REPS=2 timeout 300 $NCU -k regex:$kre -s $skip -c $cnt -o gpurun_out/$name python $PROG > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep; echo "$name done"; }
PROG=tools/ab_assembly.py
prof r02_ncu_fused_c4 k_fused_assemble 4 1 AB_CONFIG=c4
prof r02_ncu_rowfill "k_row_fill|k_row_counts" 0 2 AB_CONFIG=c4
