TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29511 tools/dist_check.py > gpurun_out/r2B_check8.log 2>&1; tail -1 gpurun_out/r2B_check8.log
for g in 8 4 2; do timeout 240 $TR --nproc-per-node $g --master-port 2952$g bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/r2B_bench$g.json 2> gpurun_out/r2B_bench$g.err; tail -1 gpurun_out/r2B_bench$g.err; done
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2B_bench1.json 2> gpurun_out/r2B_bench1.err
timeout 240 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --config trigger32768 > gpurun_out/r2B_trig8.json 2> gpurun_out/r2B_trig8.err; tail -1 gpurun_out/r2B_trig8.err
timeout 240 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 1 --no-e2e --config w16384 > gpurun_out/r2B_w8.json 2> gpurun_out/r2B_w8.err; tail -1 gpurun_out/r2B_w8.err
timeout 240 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --nmax 32768 > gpurun_out/r2B_strong8.json 2> gpurun_out/r2B_strong8.err; tail -1 gpurun_out/r2B_strong8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29534 tools/smooth_stress.py > gpurun_out/r2B_stress8.log 2>&1; tail -8 gpurun_out/r2B_stress8.log
MG_DIST_TRACE=1 timeout 200 $TR --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 2 --warmup 1 --no-e2e > /dev/null 2> gpurun_out/r2B_trace8.err; grep "mg trace" gpurun_out/r2B_trace8.err | tail -14
