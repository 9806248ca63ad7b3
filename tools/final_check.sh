#!/bin/bash
# Round-end self check on one B200 box: the commands the driver runs (GPU tests, smoke, both bench arms).
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/fin_tests.log 2>&1; echo "rc=$?" >> $O/fin_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/fin_smoke.log 2>&1; echo "rc=$?" >> $O/fin_smoke.log
timeout 600 python bench.py > $O/fin_bench.json 2> $O/fin_bench.err; echo "rc=$?" >> $O/fin_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/fin_bench_ref.json 2> $O/fin_bench_ref.err; echo "rc=$?" >> $O/fin_bench_ref.err
tail -3 $O/fin_tests.log; tail -3 $O/fin_smoke.log; cut -c1-400 $O/fin_bench.json; tail -1 $O/fin_bench.err; cut -c1-600 $O/fin_bench_ref.json; tail -1 $O/fin_bench_ref.err
