# 2-GPU check of a round: slab result of every rank == single-GPU result, weak-scaling bench line, two BASELINE configs on slabs
O=gpurun_out/${1:-r2K}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 2 --master-port 29611 tools/dist_check.py > ${O}_check2.log 2>&1; tail -1 ${O}_check2.log
timeout 240 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > ${O}_bench2.json 2> ${O}_bench2.err; tail -c 300 ${O}_bench2.json
timeout 200 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --config trigger16384 > ${O}_trig2.json 2> ${O}_trig2.err
timeout 200 $TR --nproc-per-node 2 --master-port 29614 bench.py --gpus 2 --steps 3 --warmup 1 --no-e2e --config w16384 > ${O}_w2.json 2> ${O}_w2.err
MG_DIST_TRACE=1 timeout 200 $TR --nproc-per-node 2 --master-port 29615 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e > /dev/null 2> ${O}_trace2.err; grep "mg trace" ${O}_trace2.err | tail -14
