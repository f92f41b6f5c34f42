#!/bin/bash
# Profiling recipe of one round, to be run on the GPU box through gpurun (one GPU):
#   bash scripts/profile_round.sh r01
# 1. the bench command runs plainly and must exit 0;  2. the launch list of the same command
# (gpu__time_duration, --clock-control none);  3. one --set full capture of the headline kernels with source;
# 4. one --set full capture of the TPE kernel.  Outputs land in gpurun_out/; scripts/ncu_summary.py turns them
# into the tracked summaries under profiles/.
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
cmd="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$cmd > $out/${tag}_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 $out/${tag}_plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hb_sim_kernel|kde_partial_kernel|kde_finish_kernel' -s 6 -c 2 -f \
    -o $out/${tag}_prof_bench $cmd > $out/${tag}_ncu_full.log 2>&1
# SASS-counted FP32 work of one simulation launch (not part of --set full)
ncu --clock-control none -k regex:hb_sim_kernel -s 6 -c 1 --csv --log-file $out/${tag}_sass_flops.csv --metrics \
smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_xu.sum \
    $cmd > $out/${tag}_ncu_sass.log 2>&1
python scripts/tpe_timing.py 8000 > $out/${tag}_plain_tpe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tpe_sim_kernel -s 9 -c 1 -f -o $out/${tag}_prof_tpe \
    python scripts/tpe_timing.py 8000 > $out/${tag}_ncu_tpe.log 2>&1
# early-stopping kernel (the product default), closest-known-function kernels, and the three simulation variants
python scripts/lazy_timing.py 100000 fam6 > $out/${tag}_plain_lazy.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_lazy_kernel -s 60 -c 1 -f -o $out/${tag}_prof_lazy \
    python scripts/lazy_timing.py 100000 fam6 > $out/${tag}_ncu_lazy.log 2>&1
python scripts/knownfn_timing.py 10000 > $out/${tag}_plain_knownfn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'hb_knownfn_kernel' -s 4 -c 1 -f -o $out/${tag}_prof_knownfn \
    python scripts/knownfn_timing.py 10000 > $out/${tag}_ncu_knownfn.log 2>&1
python scripts/quick_timing.py > $out/${tag}_plain_quick.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_sim_kernel -s 35 -c 1 -f -o $out/${tag}_prof_hb_flat \
    python scripts/quick_timing.py > $out/${tag}_ncu_quick.log 2>&1
tail -2 $out/${tag}_plain_bench.log | cut -c1-400
ls -la $out | grep ${tag}_
# the stand-alone nearest-recorded-arm scan (its launches come after the Hyperband ones in the same script)
ncu --set full --clock-control none --import-source on -k regex:nn_lookup_kernel -s 2 -c 1 -f -o $out/${tag}_prof_nnlookup \
    python scripts/knownfn_timing.py 10000 > $out/${tag}_ncu_nnlookup.log 2>&1
du -sh $out
