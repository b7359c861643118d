# final 1-GPU measurement of a round: smoke, bench line, reference arm, BASELINE configs, node trace, launch list
O=gpurun_out/${1:-r2J}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench1.json 2> ${O}_bench1.err; tail -c 600 ${O}_bench1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > ${O}_ref.json 2> ${O}_ref.err; tail -c 400 ${O}_ref.json
for c in v8192 w16384 trigger16384 trigger32768; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 > ${O}_cfg_$c.json 2> ${O}_cfg_$c.err; done
MG_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> ${O}_trace1.err
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > ${O}_ncu.log 2>&1
tail -1 ${O}_ncu.log
