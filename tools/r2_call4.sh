# 2-GPU run: distributed check + bench line.   gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_call4.sh'
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_solve_check.py > gpurun_out/r2d_dist2.log 2>&1; echo "dist rc=$?"; tail -4 gpurun_out/r2d_dist2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench2.json 2> gpurun_out/r2d_bench2.err; echo "bench2 rc=$?"
grep '^{' gpurun_out/r2d_bench2.json | python profiles/pick.py
