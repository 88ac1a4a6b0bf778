#!/usr/bin/env bash
# profiles/tools/slab_ncu.sh <world> <side> <out.csv> — per-launch counters of the solver kernels of ONE slab step, every rank a
# context of one process on ONE GPU (LOCAL transport; ncu serialises the launches, so each rank's kernels are timed alone)
set -u
W=$1; SIDE=$2; OUT=$3
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,dram__bytes_read.sum,dram__bytes_write.sum
SKIP=$(( 40 * W * 7 ))   # 40 settle steps x W ranks x 7-8 lambda/delta launches: the window below covers at least one step
ncu --metrics $M --clock-control none -k regex:'lambda_list|delta_list' -s $SKIP -c $(( W * 10 )) --csv --log-file $OUT python profiles/slab_local.py $W $SIDE 42 1 > /dev/null 2>&1
wc -l $OUT
