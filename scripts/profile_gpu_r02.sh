#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): every capture runs the same command once WITHOUT ncu first (exit 0), then under
# ncu.  Launch list of the bench command, then one --set full capture per kernel family.  Run under gpurun; outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-c3"
python bench.py $ARGS > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
ARGS2="--steps 1 --warmup 0 --no-cpu-baseline --no-c3"
python bench.py $ARGS2 > gpurun_out/r02_plain_bench2.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:k_full -c 2 -f -o gpurun_out/r02_prof_full python bench.py $ARGS2 > gpurun_out/r02_ncu_full.log 2>&1; echo "k_full rc=$?"
[ "${PROF_ONLY_FULL:-}" = "1" ] && exit 0      # re-capture of the headline kernel only (2: + the commit kernels)
ncu --set full --clock-control none --import-source on -k regex:"k_occ_|k_tiles" -c 6 -f -o gpurun_out/r02_prof_k2 python bench.py $ARGS2 > gpurun_out/r02_ncu_k2.log 2>&1; echo "k2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_sp_|k_str_static" -c 16 -f -o gpurun_out/r02_prof_commit python bench.py $ARGS2 > gpurun_out/r02_ncu_commit.log 2>&1; echo "commit rc=$?"
[ "${PROF_ONLY_FULL:-}" = "2" ] && exit 0
ncu --set full --clock-control none --import-source on -k regex:"k_ing_" -c 6 -f -o gpurun_out/r02_prof_ingest python bench.py $ARGS2 > gpurun_out/r02_ncu_ingest.log 2>&1; echo "ingest rc=$?"
}
# BVE evaluation kernels (K5'/K6') on a community k-SAT instance with BVE, and K9 on the C2 formula with -bce
python scripts/bve_prof.py 300000 > gpurun_out/r02_plain_bve.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bve_eval|k_bve_stream" -c 6 -f -o gpurun_out/r02_prof_bve python scripts/bve_prof.py 300000 > gpurun_out/r02_ncu_bve.log 2>&1
echo "bve rc=$?"
python scripts/bce_point.py 1000000 > gpurun_out/r02_plain_bce.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bce_masks" -c 1 -f -o gpurun_out/r02_prof_bce python scripts/bce_point.py 1000000 > gpurun_out/r02_ncu_bce.log 2>&1
echo "bce rc=$?"
ls -la gpurun_out/ | grep r02_
