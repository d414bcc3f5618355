#!/bin/bash
# One GPU-box visit of round 2: tests, tables, bench (ours + reference arm), ncu captures.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; tail -4 $O/r02_pytest_gpu.log
timeout 300 python tools/odd_sizes.py 20 > $O/r02_odd_sizes.jsonl 2> $O/odd.err; cut -c1-200 $O/r02_odd_sizes.jsonl
timeout 200 python tools/tma_check.py 512 512 256 31 31 41 20 > $O/r02_tma_c3.json 2> $O/tma.err; python tools/show_ab.py $O/r02_tma_c3.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/bench.err; tail -c 1500 $O/r02_bench.json; tail -3 $O/bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $O/r02_bench_reference.json 2> $O/bench_ref.err; tail -c 600 $O/r02_bench_reference.json
FCB200_OTF_INPLACE=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:col_tma -s 9 -c 3 -o $O/r02_tma python tools/profile_step.py 5 > $O/ncu1.log 2>&1; tail -2 $O/ncu1.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:col_tma -s 10 -c 1 -o $O/r02_tma_otf python tools/profile_step.py 5 > $O/ncu2.log 2>&1; tail -2 $O/ncu2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 8 --csv --log-file $O/r02_launches_ours.csv python tools/profile_step.py 4 > $O/ncu3.log 2>&1; tail -9 $O/r02_launches_ours.csv | cut -c1-220
ls -la $O/*.ncu-rep
