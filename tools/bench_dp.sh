#!/bin/bash
# bench lines at N GPUs of one node (what the driver launches): tools/bench_dp.sh N [workloads...]
N=$1; shift
for w in "${@:-c100}"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w \
    > gpurun_out/r02_bench_${w}_dp$N.json 2> gpurun_out/r02_bench_${w}_dp$N.err
  tail -c 600 gpurun_out/r02_bench_${w}_dp$N.json | head -c 400; echo
done
