# usage: bash scratch/cycle.sh <tag> ; runs the GPU parity tests then the 8-sequence bench
tag=$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --seqs 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c_$tag.json 2> gpurun_out/c_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
