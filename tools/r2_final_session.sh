#!/bin/bash
# Final single-GPU validation of the committed tree: GPU suite, smoke(), the reference arm and our arm of bench.py as the
# driver runs them, and the ncu launch list of the bench command.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_gpu_tests_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_smoke_final.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_reference_arm_final.json 2>$O/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2_bench_reference_arm_final.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_n1_final.json 2>$O/r2_bench_n1.err; echo "bench rc=$?"; cat $O/r2_bench_n1_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/r2_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2_launches_bench_final.out 2>&1; echo "launch list rc=$?"
timeout 900 python tools/rate_table.py > $O/r2_rate_table_final.md 2>$O/r2_rate_table_final.err; echo "rate table rc=$?"
