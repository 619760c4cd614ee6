# round 2, session zl: (1) cells kernel persistent vs one tile per CTA on C2 shuffled; (2) host-batch pipeline chunk / slot sweep
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]; e=d["e2e"]
print("  step %.4f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.4f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | e2e %.2f ms ceiling %.2f frac %.3f u32 %.2f | mismatches %d" % (e["ms_per_step"], e["copy_ceiling_ms"], e["frac_of_copy_ceiling"], e["u32_counts"]["ms_per_step"], d["parity"]["mismatches"]))'
for p in 1 0; do echo "== SIB_QC_PERSIST=$p"; SIB_QC_PERSIST=$p timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 2>/dev/null | tail -n 1 | python -c "$show"; done
for cfg in "2097152 4" "4194304 4" "8388608 4" "8388608 2" "16777216 3"; do set -- $cfg
  echo "== chunk $1 slots $2"; SIB_PIPE_CHUNK=$1 SIB_PIPE_SLOTS=$2 timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 5 --steps 5 2>/dev/null | tail -n 1 | python -c "$show"; done
