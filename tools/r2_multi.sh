#!/bin/bash
# N-GPU: the 2-rank parity tests, then the bench line at N ranks
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_multi_tests.log 2>&1
echo "multi tests exit $?"; tail -4 gpurun_out/r02_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench exit $?"; grep "bench " gpurun_out/r02_bench_n$N.err | cut -c1-220; tail -3 gpurun_out/r02_bench_n$N.err | cut -c1-300
