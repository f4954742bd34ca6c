#!/bin/bash
# Evidence pass of round 2 on one B200 (run through gpurun): the default bench line, the ncu launch list of a short bench run, and
# ncu --set full captures of the hot kernels, summarised ON the box (scripts/ncu_summary.py) so that only text travels back.
set -x
O=gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit --no-c4 --no-variants > $O/r2_final_ncu_bench.log 2>&1; echo "launchlist rc=$?"
python scripts/launch_summary.py $O/r2_final_launches.csv > $O/r2_final_launch_summary.txt
for k in k_task_objective3 k_task_optimise3 k_task_inverse3; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o /tmp/r2_$k python scripts/prof_step.py > $O/r2_final_ncu_$k.log 2>&1; echo "$k rc=$?"
  python scripts/ncu_summary.py /tmp/r2_$k.ncu-rep 30 > $O/r2_final_${k}_ncu.txt; rm -f /tmp/r2_$k.ncu-rep
done
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k_big_gemm -s 2 -c 1 -f -o /tmp/r2_k_big_gemm python scripts/bench_dense.py 4096 > $O/r2_final_ncu_big_gemm.log 2>&1; echo "biggemm rc=$?"
python scripts/ncu_summary.py /tmp/r2_k_big_gemm.ncu-rep 20 > $O/r2_final_k_big_gemm_ncu.txt; rm -f /tmp/r2_k_big_gemm.ncu-rep
for l in cpasync bulk; do echo loader=$l; MB_BIG_GEMM_LOADER=$l timeout 200 python scripts/bench_dense.py 2048 4096 8192 2>&1 | tail -9; done > $O/r2_big_gemm_loaders.txt
ls -la $O | tail -14
