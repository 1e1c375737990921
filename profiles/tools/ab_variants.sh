timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in $VARIANTS; do
FH264_B200_LIB=$PWD/scratch/libs/$v.so timeout 300 python bench.py --seqs 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if k in ('k_stage3','k_stage2','k_phase_b')})
PY
done
