#!/bin/bash
# multicast (16-byte xbar stores) against unicast pushes and against no pushes at all (FOLP_DEBUG_FLAGS=1:
# timing only, the results are wrong) on N GPUs
N=${1:-8}
mkdir -p gpurun_out
B=firstorderlp.jl_b200/libfolp_b200.so
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR --nproc-per-node $N tools/probe_kernels.py --workload target --iters 2000 $B $B:FOLP_NO_MULTICAST=1 $B:FOLP_DEBUG_FLAGS=1 > gpurun_out/probe_target_mc2_x$N.log 2> gpurun_out/probe_target_mc2_x$N.err
echo "probe target x$N rc=$?"; grep '"rank": 0' gpurun_out/probe_target_mc2_x$N.log
timeout 900 $TR --nproc-per-node $N tools/probe_kernels.py --workload c2 --iters 2000 $B $B:FOLP_NO_MULTICAST=1 > gpurun_out/probe_c2_mc2_x$N.log 2> gpurun_out/probe_c2_mc2_x$N.err
echo "probe c2 x$N rc=$?"; grep '"rank": 0' gpurun_out/probe_c2_mc2_x$N.log
