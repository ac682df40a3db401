#!/bin/bash
# round 2, multicast exchange region on N GPUs: partitioned parity in both launch forms with the
# multicast region and without it, then per-phase timings of both on the bench workload
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
FOLP_TIMING=1 timeout 600 $TR tests/dist_worker.py > gpurun_out/dist_worker_mc_x$N.log 2>&1
echo "torchrun default rc=$?"; grep -E "exchange mode|ALL OK|FAILED|no multicast" gpurun_out/dist_worker_mc_x$N.log | head
FOLP_TIMING=1 timeout 600 python tests/dist_worker.py --single-process $N > gpurun_out/multi_worker_mc_x$N.log 2>&1
echo "single-process default rc=$?"; grep -E "exchange mode|ALL OK|FAILED|no multicast" gpurun_out/multi_worker_mc_x$N.log | head
FOLP_NO_MULTICAST=1 FOLP_EXPECT_EXCHANGE=peer timeout 600 $TR tests/dist_worker.py > gpurun_out/dist_worker_nomc_x$N.log 2>&1
echo "torchrun no-multicast rc=$?"; grep -E "exchange mode|ALL OK|FAILED" gpurun_out/dist_worker_nomc_x$N.log | head
FOLP_NO_MULTICAST=1 FOLP_EXPECT_EXCHANGE=peer timeout 600 python tests/dist_worker.py --single-process $N > gpurun_out/multi_worker_nomc_x$N.log 2>&1
echo "single-process no-multicast rc=$?"; grep -E "exchange mode|ALL OK|FAILED" gpurun_out/multi_worker_nomc_x$N.log | head
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k timeout > gpurun_out/pytest_gpu_dist_timeout_x$N.log 2>&1
echo "timeout test rc=$?"; tail -2 gpurun_out/pytest_gpu_dist_timeout_x$N.log
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2}; do
  timeout 900 $TR tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_NO_MULTICAST=1 $B > gpurun_out/probe_${w}_mc_x$N.log 2> gpurun_out/probe_${w}_mc_x$N.err
  echo "probe $w rc=$?"; grep '"rank": 0' gpurun_out/probe_${w}_mc_x$N.log
done
