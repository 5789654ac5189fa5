o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/r2v_tests.log
python bench.py > $o/r2v_bench.json 2> $o/r2v_bench.err; echo "bench rc=$?"; tail -3 $o/r2v_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/r2v_bench_ref.json 2> $o/r2v_bench_ref.err; echo "ref rc=$?"; tail -c 600 $o/r2v_bench_ref.json
