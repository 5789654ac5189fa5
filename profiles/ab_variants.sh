#!/bin/bash
# A/B of build options on one B200: builds one library per variant HERE (nvcc cross-compiles without a GPU), then
#   gpurun -- 'bash profiles/ab_variants.sh run'
# times each with bench.py (device-resident leg only) and prints the per-stage milliseconds, and runs the parity tests
# on the variants that ask for it.  Variants: "name|nvcc flags|test" (test = 1: run tests/test_gpu_parity.py with it).
#   bash profiles/ab_variants.sh build      # in the authoring container
#   bash profiles/ab_variants.sh clean
VARIANTS=(
  "base||0"
)
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/split-seq_demultiplexing_b200/csrc
case "$1" in
build)
  for v in "${VARIANTS[@]}"; do
    IFS='|' read -r name flags t <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags -o "$root/profiles/v_$name.so" "$src/ssd_b200.cu" &
  done
  wait; ls -la "$root"/profiles/v_*.so ;;
run)
  cd "$root"
  for v in "${VARIANTS[@]}"; do
    IFS='|' read -r name flags t <<< "$v"
    SSD_B200_LIB=$root/profiles/v_$name.so python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-configs --no-e2e-files --no-parity-check 2>&1 | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(l['ms_per_step'], 3), {k: round(v, 3) for k, v in l['roofline']['stage_ms_per_step'].items()})"
    if [ "$t" = 1 ]; then SSD_B200_LIB=$root/profiles/v_$name.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size" 2>&1 | tail -1; fi
  done ;;
clean) rm -f "$root"/profiles/v_*.so ;;
*) echo "usage: $0 build|run|clean" ;;
esac
