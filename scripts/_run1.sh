set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01c.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_r01c.log
timeout 240 python scripts/ab_probe.py "X=1" "TSIM_STATIC_UNITS=1" "TSIM_CHUNK_ROWS=8192" "TSIM_CHUNK_ROWS=4096" > gpurun_out/ab_claim.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_claim.log
K=100 ROWS=1250000 timeout 120 python scripts/ab_probe.py "X=1" "TSIM_STATIC_UNITS=1" > gpurun_out/ab_claim_k100.log 2>&1
cat gpurun_out/ab_claim_k100.log
for v in 0 1; do
TSIM_STATIC_UNITS=$v timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:search_tc --csv --log-file gpurun_out/dram_claim_static$v.csv python scripts/profile_search.py --rows 10000000 --reps 1 > gpurun_out/dram_claim_static$v.log 2>&1
done
grep -h "dram__bytes_read\|gpu__time" gpurun_out/dram_claim_static*.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
