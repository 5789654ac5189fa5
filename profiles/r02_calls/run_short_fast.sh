# One gpurun call for the SSD_SCAN_SHORT_FAST variant (profiles/v_short.so, built here with -DSSD_SCAN_SHORT_FAST=1) against the
# shipped library: parity first, then BASELINE configs 3 and 2 side by side, then the 10 M-pair config-3 slice against the oracle.
#   gpurun --timeout 280 -- 'bash profiles/run_short_fast.sh'
o=gpurun_out; mkdir -p $o
V=/root/repo/profiles/v_short.so
B="--warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-configs --no-e2e-files --no-parity-check"
line() { python -c "
import sys, json
l = json.loads([x for x in sys.stdin.read().strip().splitlines() if x.startswith('{')][-1])
print('$1', round(l['ms_per_step'], 3), {k: round(v, 3) for k, v in l['roofline']['stage_ms_per_step'].items()}, l['result'])"; }
date +%T
SSD_B200_LIB=$V timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size" > $o/short_parity.log 2>&1; echo "parity rc=$?"; tail -2 $o/short_parity.log
date +%T
timeout 60 python bench.py --config 3 --steps 3 $B 2>$o/short_c3_base.err | line c3_base | tee $o/short_c3_base.log
SSD_B200_LIB=$V timeout 60 python bench.py --config 3 --steps 3 $B 2>$o/short_c3_var.err | line c3_short | tee $o/short_c3_var.log
date +%T
timeout 40 python bench.py --config 2 --steps 30 $B 2>$o/short_c2_base.err | line c2_base | tee $o/short_c2_base.log
SSD_B200_LIB=$V timeout 40 python bench.py --config 2 --steps 30 $B 2>$o/short_c2_var.err | line c2_short | tee $o/short_c2_var.log
date +%T
SSD_B200_LIB=$V timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_against_oracle" > $o/short_full.log 2>&1; echo "full-size vs oracle rc=$?"; tail -2 $o/short_full.log
date +%T
SSD_B200_LIB=$V timeout 100 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q > $o/short_dropin.log 2>&1; echo "dropin rc=$?"; tail -2 $o/short_dropin.log
date +%T
