# 8-GPU check of a round: slab result of every rank == single-GPU result, weak-scaling bench lines at 8 and 4 GPUs
O=gpurun_out/${1:-r2M}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29511 tools/dist_check.py > ${O}_check8.log 2>&1; tail -1 ${O}_check8.log
for g in 8 4; do timeout 240 $TR --nproc-per-node $g --master-port 2952$g bench.py --gpus $g --steps 20 --warmup 5 > ${O}_bench$g.json 2> ${O}_bench$g.err; tail -c 200 ${O}_bench$g.json; done
