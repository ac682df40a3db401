#!/bin/bash
# round 2: partitioned mode on N GPUs (N = $1, default 2): parity worker, per-phase probe of the persistent
# kernel against the kernel-per-phase form, bench line
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR tests/dist_worker.py > gpurun_out/dist_worker_x$N.log 2>&1
echo "dist_worker x$N rc=$?"; grep -E "ALL OK|FAILED|Error|error" gpurun_out/dist_worker_x$N.log | head -10; tail -3 gpurun_out/dist_worker_x$N.log
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2}; do
  timeout 600 $TR tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_PERSISTENT=0 > gpurun_out/probe_${w}_x$N.log 2> gpurun_out/probe_${w}_x$N.err
  echo "probe $w x$N rc=$?"; grep '^{' gpurun_out/probe_${w}_x$N.log | grep '"rank": 0' | cut -c1-400
done
if [ -n "$BENCH" ]; then
  timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_c2_x$N.json 2> gpurun_out/bench_c2_x$N.err
  echo "bench x$N rc=$?"; cut -c1-700 gpurun_out/bench_c2_x$N.json; grep -E "parity|PARITY|folp_create" gpurun_out/bench_c2_x$N.err | head
fi
