# Round-end measurement, part 2 (one GPU): `ncu --set full` of one step's kernels at the bench's launch size (8 sequences).
set -x
timeout 300 python bench.py --seqs 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_plain8.json 2> gpurun_out/final_plain8.err && \
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_stage|k_phase|k_features|k_interp|k_tile_index|k_scene_sad" --launch-skip 48 --launch-count 9 -f -o gpurun_out/final_full python bench.py --seqs 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu2.log 2>&1
tail -3 gpurun_out/final_ncu2.log
