#!/bin/bash
# One single-GPU gpurun session of round 2 for profiles/: launch list of the bench command, ncu --set full of the graded
# kernel (digits -1) and of its digits-11 variant, launch list of smoke().  Numbers printed by runs under ncu are not
# bench values.
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_launches_bench.out 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_pairs_const -s 100 -c 2 -f -o $O/r2_prof_const $B > $O/r2_prof_const.out 2>&1; echo "set full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_pairs_const -s 100 -c 2 -f -o $O/r2_prof_const_d11 $B --digits 11 > $O/r2_prof_const_d11.out 2>&1; echo "set full d11 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_launches_smoke.out 2>&1; echo "smoke list rc=$?"
# gpurun brings back at most 64 MiB: keep the reports' CSV pages (raw metrics, SASS/source hot spots) and one report
for r in r2_prof_const r2_prof_const_d11; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_source.csv 2>/dev/null
done
rm -f $O/r2_prof_const_d11.ncu-rep
ls -la $O/
