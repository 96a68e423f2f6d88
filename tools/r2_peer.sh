#!/bin/bash
# N GPUs: the peer-memory box reduce -- 2-rank parity test, then also.homog200 with it
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x -k "shipped_regime" -rs > gpurun_out/r02_peer_tests.log 2>&1
echo "peer tests exit $?"; tail -6 gpurun_out/r02_peer_tests.log
fi
TAMC_BENCH_PEER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 1 --packets 8000000 --no-e2e --no-cpu-baseline --also homog200 > gpurun_out/r02_peer_n$N.json 2> gpurun_out/r02_peer_n$N.err
echo "bench exit $?"; tail -3 gpurun_out/r02_peer_n$N.err | cut -c1-300
python - <<PY
import json
b=json.loads(open("gpurun_out/r02_peer_n$N.json").read().splitlines()[-1])
h=b["also"]["homog200"]
print({k:round(h[k],3) if isinstance(h[k],float) else h[k] for k in ("kernel_ms","ms_per_step","allreduce_ms")}, [round(x,3) for x in h["kernel_ms_by_rank"]])
print(h.get("peer_reduce"))
PY
