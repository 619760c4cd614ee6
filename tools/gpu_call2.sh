# round-1 session-5 GPU call: parity tests, bench of the rank-cells path, ncu evidence (r01d)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log; tail -5 gpurun_out/tests.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 4000 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d" % d["parity"]["mismatches"], d["config"].get("rank_cells"))'
echo "== cells through the partition"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 --partition 2>&1 | tail -1 | python -c "$show" > gpurun_out/cells_partition.log 2>&1; cat gpurun_out/cells_partition.log
echo "== sorted"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 --order sorted 2>&1 | tail -1 | python -c "$show" > gpurun_out/sorted.log 2>&1; cat gpurun_out/sorted.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_count_cells" -s 2 -c 1 -o gpurun_out/prof_cells_c2_r01d -f python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_count_cells" -s 1 -c 1 -o gpurun_out/prof_cells_c2sorted_r01d -f python tools/prof_driver.py c2 count 2 > gpurun_out/p2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_scan|qk_fill|qk_count" -c 6 -o gpurun_out/prof_search_c3_r01d -f python tools/prof_driver.py c3 search 1 > gpurun_out/p3.log 2>&1
ls -la gpurun_out
