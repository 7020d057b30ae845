#!/bin/bash
# Round-2 evidence: in-graph timelines, ncu launch list of one eager step, ncu --set full of the roofline kernel and of the tensor-core kernels.
mkdir -p gpurun_out
timeout 200 python tools/timeline.py --raw --cta-iter 6 > gpurun_out/r02_timeline.txt 2>&1
timeout 200 python tools/timeline.py --precision f16 --raw > gpurun_out/r02_timeline_f16.txt 2>&1
timeout 200 python tools/timeline.py --correlation otf > gpurun_out/r02_timeline_otf.txt 2>&1
timeout 200 python tools/timeline.py --preset E_I_LU5_BD10 --batch 4 --h 384 --w 512 > gpurun_out/r02_timeline_configM_b4.txt 2>&1
timeout 200 python tools/timeline.py --batch 4 > gpurun_out/r02_timeline_configD_b4.txt 2>&1
timeout 200 python tools/step_profile.py --all > gpurun_out/r02_step_profile.txt 2>&1
N=$(python -c "import re;print(re.search(r'of (\d+);', open('gpurun_out/r02_timeline.txt').readline()).group(1))")
echo "launches per forward: $N"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:bflow:: --launch-skip $N -c $N --csv --log-file gpurun_out/r02_launches_step.csv python tools/one_step.py > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:corr_lookup_tiled --launch-skip 14 -c 1 -f -o gpurun_out/r02_lookup_instep python tools/one_step.py > gpurun_out/ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_lookup_instep.ncu-rep > gpurun_out/r02_lookup_instep.txt 2>&1; rm -f gpurun_out/r02_lookup_instep.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_tc3_kernel --launch-skip 205 -c 9 -f -o gpurun_out/r02_tc3_update python tools/one_step.py > gpurun_out/ncu3.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_tc3_update.ncu-rep > gpurun_out/r02_tc3_update.txt 2>&1; rm -f gpurun_out/r02_tc3_update.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_slab64 --launch-skip 4 -c 1 -f -o gpurun_out/r02_slab64 python tools/one_step.py > gpurun_out/ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_slab64.ncu-rep > gpurun_out/r02_slab64.txt 2>&1; rm -f gpurun_out/r02_slab64.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:instnorm_relu16_v8 --launch-skip 13 -c 2 -f -o gpurun_out/r02_instnorm python tools/one_step.py > gpurun_out/ncu5.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_instnorm.ncu-rep > gpurun_out/r02_instnorm.txt 2>&1; rm -f gpurun_out/r02_instnorm.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:corr_lookup_tiled -c 1 -f -o gpurun_out/r02_lookup_b32 python tools/lookup_bench.py --batch 32 --tiled --reps 2 > gpurun_out/ncu6.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_lookup_b32.ncu-rep > gpurun_out/r02_lookup_b32.txt 2>&1; rm -f gpurun_out/r02_lookup_b32.ncu-rep
ls -la gpurun_out | grep r02_
