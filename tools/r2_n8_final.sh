#!/bin/bash
# 8 GPUs: (1) the homog200 kernel in eight independent processes, no communicator; (2) the bench line of the round's last state
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pids=()
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i timeout 120 python tools/dbg_homog_solo.py 8 > gpurun_out/solo_$i.log 2>&1 &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
cat gpurun_out/solo_*.log | tee gpurun_out/r02_n8_solo.txt
CUDA_VISIBLE_DEVICES=3 timeout 120 python tools/dbg_homog_solo.py 4 | tee -a gpurun_out/r02_n8_solo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "bench exit $?"; grep "bench " gpurun_out/r02_bench_n8.err | cut -c1-200 | head -14
python - <<PY
import json
b=json.loads(open("gpurun_out/r02_bench_n8.json").read().splitlines()[-1])
print(b["value"], b["ms_per_step"], b["e2e"]["value"], (b["e2e"].get("root_io") or {}).get("value"))
h=b["also"]["homog200"]
print({k:h.get(k) for k in ("kernel","kernel_ms","ms_per_step","allreduce_ms","e2e_ms_per_step","e2e_root_io_ms_per_step")})
print("by rank", [round(x,3) for x in h["kernel_ms_by_rank"]])
c=b["also"]["coupled_calls_shipped80"]; print(c["mean_us"], c["p95_us"], c.get("root_io",{}).get("mean_us"))
PY
