# Development helper (N GPUs): distributed parity check, then the bench with the CG phase trace.
#   gpurun --gpus N --timeout 900 -- 'bash tools/dist2.sh N'
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/dist_solve_check.py 12 2>&1 | grep -E "^\[dist|DIST|rror" | tail -14
FDB_CG_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N ${BENCH_ARGS:-} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -E "trace|rror" gpurun_out/bench_n$N.err | tail -6
python profiles/pick.py < gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print("parity", json.dumps(d.get("parity")))
c3=d.get("configs",{}).get("c3",{})
print("c3 solve", json.dumps(c3.get("solve")), json.dumps(c3.get("parity")), c3.get("error"))
PY
