#!/bin/bash
# Round-end validation in one gpurun call: full GPU suite, smoke, bench line, ncu launch list + one full capture per kernel.
mkdir -p gpurun_out
T=${1:-r1q}
timeout 400 python -m pytest tests -q -m gpu --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/${T}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${T}_smoke.log
timeout 400 python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${T}.json; tail -n 3 gpurun_out/bench_${T}.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py --pages 64 --steps 2 --warmup 1 > gpurun_out/${T}_ncu_l.log 2>&1; echo "launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_opt_iir|k_opt_fir|k_sauvola|k_resample|k_gray_blur_fast|k_noise|k_mask_denoise" -c 9 -o gpurun_out/${T}_full -f python tools/profile_step.py --pages 64 --warmup 0 --steps 1 > gpurun_out/${T}_ncu_f.log 2>&1; echo "full rc=$?"; tail -n 2 gpurun_out/${T}_ncu_f.log
