#!/bin/bash
# partitioned mode on N GPUs: parity worker, then the probe with and without the copy-engine pushes
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR tests/dist_worker.py > gpurun_out/dist_worker_x$N.log 2>&1
echo "dist_worker x$N rc=$?"; grep -E "ALL OK|FAILED" gpurun_out/dist_worker_x$N.log | head -5; tail -2 gpurun_out/dist_worker_x$N.log
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2}; do
  timeout 900 $TR tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_DEBUG_FLAGS=8 $B:FOLP_DEBUG_FLAGS=16 $B:FOLP_DEBUG_FLAGS=24 $B > gpurun_out/probe_${w}_x$N.log 2> gpurun_out/probe_${w}_x$N.err
  echo "probe $w x$N rc=$?"; grep '^{' gpurun_out/probe_${w}_x$N.log | grep '"rank": 0' | cut -c1-330
done
