o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
