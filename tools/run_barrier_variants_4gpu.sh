P="import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value %.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'kernel', round(d['roofline']['us_per_launch'],2), 'verified', d.get('exchange_verified'), 'gave_up', d.get('barrier_gave_up'))"
for B in ${BARS:-none fused flags nccl}; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --no-cpu --no-parity --barrier $B --e2e-steps 4 > gpurun_out/b4_$B.json 2> gpurun_out/b4_$B.err; python -c "$P" gpurun_out/b4_$B.json
done
