#!/usr/bin/env bash
# profiles/tools/ab_ncu.sh <side> <variant> ... — a few L1/L2 counters of one lambda launch per variant (ncu, --metrics only)
set -u
SIDE=$1; shift
M=gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
cp pbf_sph_b200/libpbf_cuda.so /tmp/libpbf_cuda.orig.so
for v in "$@"; do
  cp exp_libs/libpbf_$v.so pbf_sph_b200/libpbf_cuda.so
  echo "== $v"
  ncu --metrics $M --clock-control none -k regex:lambda_list -s 400 -c 1 --csv python profiles/tools/size_sweep.py $SIDE 2>&1 | grep -E '^"[0-9]' | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"'
done
cp /tmp/libpbf_cuda.orig.so pbf_sph_b200/libpbf_cuda.so
