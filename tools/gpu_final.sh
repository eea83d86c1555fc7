#!/bin/bash
# Round-end validation in one gpurun call: full GPU suite, smoke, bench line, ncu launch list + one full capture per kernel.
#   gpurun --timeout 1800 -- 'bash tools/gpu_final.sh r2u'
mkdir -p gpurun_out
T=${1:-r2u}
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${T}_bench.json; tail -n 3 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${T}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_64pages.csv python tools/profile_step.py --pages 64 --steps 2 --warmup 1 > gpurun_out/${T}_ncu_l.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_noise|k_gray|k_sauvola_fused|k_mask|k_opt|k_resample}" -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-10} -o gpurun_out/${T}_full -f python tools/profile_step.py --pages 64 --warmup 1 --steps 1 > gpurun_out/${T}_ncu_f.log 2>&1; echo "full rc=$?"; tail -n 2 gpurun_out/${T}_ncu_f.log
