#!/bin/bash
# Runs on the GPU box (under gpurun): a reduced bench command plain, then the ncu launch list of the
# SAME command, then one --set full capture of each top kernel and an instruction-mix pass of the sweep kernel.
# Outputs go to gpurun_out/; tools/make_profiles.py <tag> turns them into profiles/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --fields 32 --gs-events 1024"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || { echo "plain run failed"; tail -n 5 gpurun_out/bench_plain.err; exit 1; }
[ -n "${SKIP_GS:-}" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep_bricks" -s 1 -c 1 -o gpurun_out/prof_fsm -f $CMD > gpurun_out/ncu_fsm.log 2>&1
# the search kernel's capture replays a 2.7 s launch ~40 times: skip it with SKIP_GS=1 when only the sweep kernel changed
[ -n "${SKIP_GS:-}" ] || ncu --set full --clock-control none --import-source on -k regex:"locate_uniform" -c 1 -o gpurun_out/prof_gs -f $CMD > gpurun_out/ncu_gs.log 2>&1
PIPES="smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,smsp__thread_inst_executed.sum"
ncu --metrics $PIPES --clock-control none -k regex:"sweep_bricks" -s 1 -c 1 --csv --log-file gpurun_out/pipes_fsm.csv $CMD > gpurun_out/ncu_pipes.log 2>&1
tail -n 2 gpurun_out/ncu_fsm.log gpurun_out/ncu_gs.log
