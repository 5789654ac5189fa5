# Second (last) call for SSD_SCAN_SHORT_FAST: the variant with the UMI masks uniform again (profiles/v_short2.so) against the shipped
# library on config 2, then the parity tests that reach the changed code (short read 2, N bases, ties, both scan variants).
o=gpurun_out; mkdir -p $o
V=/root/repo/profiles/v_short2.so
B="--warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-configs --no-e2e-files --no-parity-check"
line() { python -c "
import sys, json
l = json.loads([x for x in sys.stdin.read().strip().splitlines() if x.startswith('{')][-1])
print('$1', round(l['ms_per_step'], 3), {k: round(v, 3) for k, v in l['roofline']['stage_ms_per_step'].items()}, l['result'])"; }
date +%T
timeout 20 python bench.py --config 2 --steps 30 $B 2>/dev/null | line c2_base | tee $o/short2_c2_base.log
SSD_B200_LIB=$V timeout 20 python bench.py --config 2 --steps 30 $B 2>/dev/null | line c2_short2 | tee $o/short2_c2_var.log
date +%T
SSD_B200_LIB=$V timeout 45 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rule_cases or adversarial or synthetic_vs_oracle or per_round_lists_vs_oracle or cell_umi_table or wide_offsets or config3 or scan_without_guessing or misleading or short_records or fuzz or bundled_digests or record_cells" > $o/short2_parity.log 2>&1; echo "parity rc=$?"; tail -2 $o/short2_parity.log
date +%T
