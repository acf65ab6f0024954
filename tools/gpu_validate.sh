#!/bin/bash
# One gpurun call that validates everything this repo claims on ONE B200 and leaves the evidence in gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_validate.sh r02'
# (prefix = file-name prefix for the outputs; copy what should be judged into profiles/ afterwards)
set -u
P=${1:-val}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $O/${P}_tests.log 2>&1; echo "rc=$?" >> $O/${P}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1; echo "rc=$?" >> $O/${P}_smoke.log
# the plain-C consumer of the ABI
gcc -std=c99 -I include examples/hexagonal.c -L rome.jl_b200 -lrome_b200 -Wl,-rpath,$PWD/rome.jl_b200 -lm -o /tmp/hexagonal \
  && timeout 60 /tmp/hexagonal > $O/${P}_hexagonal_c.log 2>&1; echo "rc=$?" >> $O/${P}_hexagonal_c.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/${P}_bench_ref.json 2> $O/${P}_bench_ref.err
timeout 300 python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
for W in beehive_N200 se3_chain_10k; do
  timeout 300 python bench.py --workload $W --cpu-seconds 4 > $O/${P}_bench_${W}_n1.json 2> $O/${P}_bench_${W}_n1.err
done
# launch list of the same command (times under ncu are cold-cache and serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu --no-parity --e2e-steps 4 > $O/${P}_b_ncu.log 2>&1
# full captures of the hot kernels
for F in pose2pose2 bearingrange pose3pose3; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:eval_kernel -c 2 -f -o $O/${P}_$F \
    python tools/prof_kernel.py $F 1 > $O/${P}_ncu_$F.log 2>&1
done
timeout 200 python tools/bench_families.py > $O/${P}_families.jsonl 2> $O/${P}_families.err
# device-resident sweeps (convolutions + product of proposals) and a full capture of the product kernel
timeout 100 python tools/bench_solver.py > $O/${P}_solver.json 2> $O/${P}_solver.err
timeout 150 ncu --set full --import-source on --clock-control none -k regex:product_kernel -c 1 --launch-skip 1 -f -o $O/${P}_product \
  python tools/bench_solver.py > $O/${P}_ncu_product.log 2>&1
tail -2 $O/${P}_tests.log; tail -1 $O/${P}_smoke.log; tail -3 $O/${P}_hexagonal_c.log; cut -c1-300 $O/${P}_bench_n1.json
