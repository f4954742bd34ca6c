#!/bin/bash
# Second evidence pass of round 2 (after the E-step / monitoring changes): default bench line, launch list, E-step kernel capture.
set -x
O=gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit --no-c4 --no-variants > $O/r2_final_ncu_bench.log 2>&1; echo "launchlist rc=$?"
python scripts/launch_summary.py $O/r2_final_launches.csv > $O/r2_final_launch_summary.txt
for k in k_task_inverse3; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/r2_$k python scripts/prof_step.py > $O/r2_final_ncu_$k.log 2>&1; echo "$k rc=$?"
  python scripts/ncu_summary.py /tmp/r2_$k.ncu-rep 30 > $O/r2_final_${k}_ncu.txt; rm -f /tmp/r2_$k.ncu-rep
done
