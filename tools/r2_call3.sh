# Round-2 check of the faster row-wise pattern kernels.   gpurun --timeout 900 -- 'bash tools/r2_call3.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rowwise or bit_identical or operators or reproducible" > gpurun_out/r2c_quick.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/r2c_quick.log
timeout 400 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python profiles/pick.py < gpurun_out/r2c_bench.json
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_gputests.log
NCU="ncu --set full --clock-control none --import-source on -f"
prof() { name=$1; kre=$2; skip=$3; cnt=$4; shift 4; env "$@" AB_REPS=2 timeout 300 $NCU -k regex:$kre -s $skip -c $cnt -o gpurun_out/$name python $PROG > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep; echo "$name done"; }
PROG=tools/ab_assembly.py
prof r02_ncu_rowfill "k_row_fill|k_row_counts" 0 2 AB_CONFIG=c4
