o=gpurun_out
bash profiles/ab_variants.sh run > $o/r2j_ab.log 2>&1
SSD_B200_LIB=/root/repo/profiles/v_prof.so python profiles/prof_phases.py > $o/r2j_phases.log 2>&1
cat $o/r2j_ab.log $o/r2j_phases.log
