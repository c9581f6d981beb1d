# Round-2 evidence run (one GPU): GPU tests, bench (both arms), ncu launch list of the bench command, full ncu
# captures of the dominant kernels.   gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gputests.log
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --min-warmup-s 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench_ncu.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
prof() { name=$1; kre=$2; skip=$3; cnt=$4; shift 4; env "$@" AB_REPS=2 timeout 300 $NCU -k regex:$kre -s $skip -c $cnt -o gpurun_out/$name python $PROG > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details --csv > gpurun_out/${name}_details.csv 2>/dev/null; rm -f gpurun_out/$name.ncu-rep; echo "$name done"; }
PROG=tools/ab_assembly.py
prof r02_ncu_fused_c4 k_fused_assemble 4 1 AB_CONFIG=c4
prof r02_ncu_fused_c2 k_fused_assemble 4 1 AB_CONFIG=c2
prof r02_ncu_fused_c3 k_fused_assemble 4 1 AB_CONFIG=c3
prof r02_ncu_fused_p2tet k_fused_assemble 4 1 AB_CONFIG=p2tet
PROG=tools/solver_ab.py
prof r02_ncu_cg "k_spmv_sell|k_cg_" 30 4 AB_SOLVERS=c4only
du -sh gpurun_out; ls -la gpurun_out | head -40
