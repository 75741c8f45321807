#!/bin/bash
# Round-end ncu evidence (run under gpurun on ONE GPU; numbers printed by runs under ncu are never bench values).
#   1. launch list of the bench command (per-launch gpu__time_duration, cold and serialised: shares, not absolutes)
#   2. --set full of the dominant kernel (RIS candidates) for the roofline `traffic` field
#   3. --set full of the visibility trace kernel and of the NRC inference kernels (grid encode with the TMA-staged level
#      table, tcgen05 MLP) and the tcgen05 training kernel
# Outputs land in gpurun_out/; summaries are copied to profiles/ by hand (tools/ncu_pick.py prints the quoted metrics).
set -u
TAG=${TAG:-r02}
mkdir -p gpurun_out
timeout -s KILL 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 170 ncu --set full --clock-control none --import-source on -k regex:k_initialAndTemporalRIS -s 6 -c 1 -f -o gpurun_out/ris_${TAG} \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --headline-only > gpurun_out/ncu_ris.log 2>&1
echo "ris full rc=$?"
timeout -s KILL 170 ncu --set full --clock-control none --import-source on -k regex:k_traceWavefrontDeferred -s 4 -c 1 -f -o gpurun_out/trace_${TAG} \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --headline-only > gpurun_out/ncu_trace.log 2>&1
echo "trace full rc=$?"
timeout -s KILL 170 ncu --set full --clock-control none --import-source on -k regex:"k_nrcGridEncode|k_nrcInferMlp|k_nrcTrainTc" -s 8 -c 3 -f -o gpurun_out/nrc_${TAG} \
    python tools/nrc_bench.py > gpurun_out/ncu_nrc.log 2>&1
echo "nrc full rc=$?"
ls -la gpurun_out/*${TAG}* 2>/dev/null
