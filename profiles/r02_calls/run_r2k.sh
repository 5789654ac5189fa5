o=gpurun_out
bash profiles/ab_variants.sh run > $o/r2k_ab.log 2>&1
cat $o/r2k_ab.log
