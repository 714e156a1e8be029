#!/bin/bash
# Runs on the GPU box (under gpurun): a reduced bench command plain, then the ncu launch list of the
# SAME command, then one --set full capture of each top kernel.  Outputs go to gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --fields 32 --gs-events 64"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || { echo "plain run failed"; tail -n 5 gpurun_out/bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep_bricks" -s 1 -c 1 -o gpurun_out/prof_fsm -f $CMD > gpurun_out/ncu_fsm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"locate_uniform" -c 1 -o gpurun_out/prof_gs -f $CMD > gpurun_out/ncu_gs.log 2>&1
tail -n 2 gpurun_out/ncu_fsm.log gpurun_out/ncu_gs.log
