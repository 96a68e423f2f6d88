cd "${GRAFT_REPO_ROOT:-/root/repo}"
for opt in "reduce_bound=1" "reduce_bound=2"; do
python bench.py --no-cpu-baseline --no-also --steps 20 --option $opt > gpurun_out/ab.json 2>/dev/null
python - <<PY
import json
b=json.load(open('gpurun_out/ab.json'))
print('$opt', 'value %.4g ms %.4f breakdown %s launches %s' % (b['value'], b['ms_per_step'], b['breakdown_ms_per_step'], b['gpu_launches']))
PY
done
