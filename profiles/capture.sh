#!/usr/bin/env bash
# [WORKLOAD=dam-8m] profiles/capture.sh <tag> — run on a B200 box (under gpurun): plain bench, ncu launch list of the same command,
# one `ncu --set full` capture of the solver-iteration kernels.  Outputs land in gpurun_out/<tag>_*.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --settle 30 --no-cpu-baseline --no-e2e --no-secondary --workload ${WORKLOAD:-dam-1m}"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
# launch list: skip the settle+warm-up launches, list the 3 timed steps
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-900} -c ${COUNT:-120} --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-lambda|delta}" -s ${FSKIP:-200} -c ${FCOUNT:-2} \
    -o $OUT/${TAG}_full -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -8
