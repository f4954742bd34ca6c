#!/bin/bash
# Final evidence pass of round 2 (after the pair tables / upload / E-step reservation changes), one B200 through gpurun:
# default bench line, ncu launch list of a short bench run, ncu --set full of the dominant kernel, the E-step kernel and
# the pair-table builder (summarised on the box by scripts/ncu_summary.py), then configs 3 and 5 at full size.
set -x
O=gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2c_final_bench.json 2> $O/r2c_final_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2c_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit --no-c4 --no-variants > $O/r2c_final_ncu_bench.log 2>&1; echo "launchlist rc=$?"
python scripts/launch_summary.py $O/r2c_final_launches.csv > $O/r2c_final_launch_summary.txt
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k_task_objective3 -s 2 -c 1 -f -o /tmp/r2c_obj python scripts/prof_obj.py > $O/r2c_ncu_obj.log 2>&1; echo "obj rc=$?"
python scripts/ncu_summary.py /tmp/r2c_obj.ncu-rep 30 > $O/r2c_final_k_task_objective3_ncu.txt; rm -f /tmp/r2c_obj.ncu-rep
for k in k_task_inverse3 k_task_pairs3; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/r2c_$k python scripts/prof_step.py > $O/r2c_ncu_$k.log 2>&1; echo "$k rc=$?"
  python scripts/ncu_summary.py /tmp/r2c_$k.ncu-rep 25 > $O/r2c_final_${k}_ncu.txt; rm -f /tmp/r2c_$k.ncu-rep
done
timeout 300 python scripts/bench_c3.py > $O/r2c_c3.log 2>&1; echo "c3 rc=$?"; tail -3 $O/r2c_c3.log
timeout 300 python scripts/bench_c5.py > $O/r2c_c5.log 2>&1; echo "c5 rc=$?"; tail -3 $O/r2c_c5.log
ls -la $O | tail -16
