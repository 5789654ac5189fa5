#!/bin/bash
# Round capture on one B200: the bench line, then (only after the bench exited 0 without ncu) the ncu launch list and one
# full-set capture of the dominant kernels.  Everything lands in gpurun_out/; profiles/README.md says how it is summarised.
# usage: bash profiles/capture.sh [tag]
tag=${1:-r02}
o=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-other-configs --no-e2e-files --no-parity-check"
$B > $o/plain_$tag.log 2>&1 || { echo "plain bench failed"; tail -5 $o/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv $B > $o/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^(k_scan|k_emit_stream|k_umi_sets|k_umi_bitmap_counts)$" -s 10 -c 5 -f -o $o/top4_$tag $B > $o/ncu_f_$tag.log 2>&1
ncu -i $o/top4_$tag.ncu-rep --page raw --csv > $o/top4_${tag}_raw.csv 2>/dev/null
ncu -i $o/top4_$tag.ncu-rep --page source --print-source cuda,sass --csv > $o/top4_${tag}_src.csv 2>/dev/null
rm -f $o/top4_$tag.ncu-rep
ls -la $o/*$tag* | awk '{print $5, $9}'
