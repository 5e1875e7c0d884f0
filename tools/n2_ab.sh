T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29531 bench.py --gpus 2 --steps 10 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n2_hp.json 2> gpurun_out/r2_n2_hp.err
VU_DP_HIGH_PRIORITY=0 timeout 300 $T --master-port 29532 bench.py --gpus 2 --steps 10 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n2_nohp.json 2> gpurun_out/r2_n2_nohp.err
timeout 300 python bench.py --gpus 1 --steps 10 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n1_nt.json 2>/dev/null
timeout 300 $T --master-port 29533 tools/dp_timeline.py 256 > gpurun_out/r2_dp_timeline_n2_hp.txt 2>/dev/null
for f in r2_n2_hp r2_n2_nohp r2_n1_nt; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',round(d['value'],1),round(d['ms_per_step'],2))"; done
tail -8 gpurun_out/r2_dp_timeline_n2_hp.txt; tail -2 gpurun_out/r2_n2_hp.err
