#!/bin/bash
# multicast against unicast pushes: per-phase timings of the partitioned take_step on N GPUs (and on 4 of them)
N=${1:-8}
mkdir -p gpurun_out
B=firstorderlp.jl_b200/libfolp_b200.so
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29519"
for w in ${WORKLOADS:-c2 target}; do
  timeout 900 $TR --nproc-per-node $N tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_NO_MULTICAST=1 > gpurun_out/probe_${w}_mc_x$N.log 2> gpurun_out/probe_${w}_mc_x$N.err
  echo "probe $w x$N rc=$?"; grep '"rank": 0' gpurun_out/probe_${w}_mc_x$N.log
done
if [ "$N" -gt 4 ]; then
  timeout 600 $TR --nproc-per-node 4 tools/probe_kernels.py --workload c2 --iters 2000 $B $B:FOLP_NO_MULTICAST=1 > gpurun_out/probe_c2_mc_x4.log 2> gpurun_out/probe_c2_mc_x4.err
  echo "probe c2 x4 rc=$?"; grep '"rank": 0' gpurun_out/probe_c2_mc_x4.log
fi
