set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --small-q \"\""
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; echo "bench rc=$?"
cat gpurun_out/bench_r01c.json | cut -c1-600
timeout 200 python scripts/ab_probe.py "X=1" "TSIM_CHUNK_ROWS=32768" "TSIM_CHUNK_ROWS=65536" > gpurun_out/ab_chunk2.log 2>&1; cat gpurun_out/ab_chunk2.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:search_|select_|merge_|row_inv|tighten|pool_' --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --small-q "" > gpurun_out/ncu_launches_r01c.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 2 -c 1 -f -o gpurun_out/prof_r01c_bench_q4096 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --small-q "" > gpurun_out/ncu_full_r01c.log 2>&1; echo "ncu2 rc=$?"
timeout 200 python scripts/time_configs.py c2 c3 c5 > gpurun_out/time_configs_r01c.log 2>&1; cat gpurun_out/time_configs_r01c.log
