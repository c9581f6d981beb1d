# final defaults: tests + bench.   gpurun --timeout 1200 -- 'bash tools/r2_call14.sh'
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_gputests.log
timeout 400 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python profiles/pick.py < gpurun_out/r2h_bench.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()"
NCU="ncu --set full --clock-control none --import-source on -f"
name=r02_ncu_persist_c3
AB_REPS=2 AB_CONFIG=c3 timeout 300 $NCU -k regex:k_fused_persist -s 4 -c 1 -o gpurun_out/$name python tools/ab_assembly.py > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep; echo "ncu done"
