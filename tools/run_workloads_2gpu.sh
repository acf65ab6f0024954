# N=2 (and N=1) runs of the three bench workloads on a 2-GPU box; prints one summary line per run
P="import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], '%.4g'%d['value'], round(d['ms_per_step']*1e3,2), round(d['roofline']['us_per_launch'],2), round(d['roofline']['frac'],3), d.get('exchange_verified'), d.get('barrier_gave_up'), d.get('max_abs_diff'), d.get('parity',{}).get('ok'), d.get('parity',{}).get('max_abs_err'))"
for W in ${WORKLOADS:-manhattan_shaped_10k_se2_N100 beehive_N200 se3_chain_10k}; do
  timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-cpu --workload $W > gpurun_out/${TAG:-m}_${W}_n2.json 2> gpurun_out/${TAG:-m}_${W}_n2.err; grep -i "error\|mismatch\|Traceback" gpurun_out/${TAG:-m}_${W}_n2.err | head -3; python -c "$P" gpurun_out/${TAG:-m}_${W}_n2.json
done
for W in ${WORKLOADS1:-}; do
  timeout 250 python bench.py --no-cpu --workload $W > gpurun_out/${TAG:-m}_${W}_n1.json 2> gpurun_out/${TAG:-m}_${W}_n1.err; grep -i "error\|Traceback" gpurun_out/${TAG:-m}_${W}_n1.err | head -3; python -c "$P" gpurun_out/${TAG:-m}_${W}_n1.json
done
