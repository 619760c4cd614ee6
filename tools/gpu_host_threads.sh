# host-side copy threads for the search_values end-to-end call (pageable result buffers)
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); sv=d["search_values"]
print("  count e2e %.2f ms | search_values e2e %.2f ms (first %.1f) device %.3f ms" % (d["e2e"]["ms_per_step"], sv["e2e"]["ms_per_step"], sv["e2e"]["first_call_ms"], sv["ms_per_step"]))'
for t in 8 4 12 16; do
  echo "== SIB_HOST_THREADS=$t"; SIB_HOST_THREADS=$t timeout 300 python bench.py --steps 2 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1 | python -c "$show"
done > gpurun_out/host_threads.log 2>&1
cat gpurun_out/host_threads.log; nproc
