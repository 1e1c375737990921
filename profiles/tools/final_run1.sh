# Round-end measurement, part 1 (one GPU): parity suite, bench (own arm + reference arm), ncu launch list of the bench command.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 400 gpurun_out/final_ref.json
timeout 300 python bench.py --seqs 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --seqs 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu1.log 2>&1
tail -2 gpurun_out/final_ncu1.log
