#!/bin/bash
# One 8-GPU call: weak scaling of the default workload at N = 8 / 4 with the one-kernel flag barrier (default) and with the
# barrier fused into the evaluation kernels, the SE(3) chain on 8 GPUs, and the NVLink byte counters around the N = 8 run.
P="import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value %.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'kernel', round(d['roofline']['us_per_launch'],2), 'verified', d.get('exchange_verified'), 'gave_up', d.get('barrier_gave_up'), 'rows', d.get('rows_checked_all_ranks'), 'parity', d.get('parity',{}).get('ok'))"
run() {  # workload gpus tag extra...
  W=$1; G=$2; T=$3; shift 3
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --no-cpu --workload $W "$@" > gpurun_out/$T.json 2> gpurun_out/$T.err
  grep -i "error\|mismatch\|Traceback" gpurun_out/$T.err | head -3; python -c "$P" gpurun_out/$T.json
}
T=manhattan_shaped_10k_se2_N100
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_before.txt 2>&1
run $T 8 ${TAG}_T_n8
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_after.txt 2>&1
run $T 8 ${TAG}_T_n8_fused --barrier fused --no-parity --e2e-steps 4
run se3_chain_10k 8 ${TAG}_se3_n8 --e2e-steps 8
run $T 4 ${TAG}_T_n4 --no-parity --e2e-steps 4
