#!/bin/bash
# One gpurun call that validates everything this repo claims on a B200 box and leaves the evidence in gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_validate.sh r02'
# (prefix = file-name prefix for the outputs; copy what should be judged into profiles/ afterwards)
set -u
P=${1:-val}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $O/${P}_tests.log 2>&1; echo "rc=$?" >> $O/${P}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1; echo "rc=$?" >> $O/${P}_smoke.log
# the plain-C consumer of the ABI: not yet run on a GPU in round 1
gcc -std=c99 -I include examples/hexagonal.c -L rome.jl_b200 -lrome_b200 -Wl,-rpath,$PWD/rome.jl_b200 -lm -o /tmp/hexagonal \
  && /tmp/hexagonal > $O/${P}_hexagonal_c.log 2>&1; echo "rc=$?" >> $O/${P}_hexagonal_c.log
python bench.py --impl reference --steps 20 --warmup 3 > $O/${P}_bench_ref.json 2> $O/${P}_bench_ref.err
python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
# launch list of the same command (times under ncu are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu > $O/${P}_b_ncu.log 2>&1
tail -2 $O/${P}_tests.log; tail -1 $O/${P}_smoke.log; tail -1 $O/${P}_hexagonal_c.log; cut -c1-300 $O/${P}_bench_n1.json
