# 8-GPU box: weak scaling of the default workload at N = 8 / 4 (/ 2 / 1), and the BASELINE multi-GPU configs on the GPUs
# they name (Beehive N=200 on 4, SE(3) chain on 8).  One summary line per run; JSON lines under gpurun_out/.
P="import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value %.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'kernel', round(d['roofline']['us_per_launch'],2), 'verified', d.get('exchange_verified'), 'gave_up', d.get('barrier_gave_up'), 'rows', d.get('rows_checked_all_ranks'), 'parity', d.get('parity',{}).get('ok'))"
run() {  # workload gpus tag
  if [ "$2" = 1 ]; then
    timeout 300 python bench.py --no-cpu --workload $1 > gpurun_out/$3.json 2> gpurun_out/$3.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $2 --no-cpu --workload $1 > gpurun_out/$3.json 2> gpurun_out/$3.err
  fi
  grep -i "error\|mismatch\|Traceback" gpurun_out/$3.err | head -3; python -c "$P" gpurun_out/$3.json
}
T=manhattan_shaped_10k_se2_N100
run $T 8 ${TAG}_T_n8
run se3_chain_10k 8 ${TAG}_se3_n8
run $T 4 ${TAG}_T_n4
run beehive_N200 4 ${TAG}_beehive_n4
if [ -z "$SHORT" ]; then
run $T 2 ${TAG}_T_n2
run $T 1 ${TAG}_T_n1
fi
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
