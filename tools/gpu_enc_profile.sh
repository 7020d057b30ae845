#!/bin/bash
# ncu --set full of the encoder's tensor-core launches (one eager forward, single stream): stem, layer1 slabs, layer2/3 + output convs
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_tc3_kernel -c 36 -f -o gpurun_out/enc_tc3 python tools/one_step.py --n 1 > gpurun_out/ncu_enc1.log 2>&1
python tools/ncu_summary.py gpurun_out/enc_tc3.ncu-rep > gpurun_out/enc_tc3.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_slab64 -c 8 -f -o gpurun_out/enc_slab64 python tools/one_step.py --n 1 > gpurun_out/ncu_enc2.log 2>&1
python tools/ncu_summary.py gpurun_out/enc_slab64.ncu-rep > gpurun_out/enc_slab64.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_stem7 -c 2 -f -o gpurun_out/enc_stem7 python tools/one_step.py --n 1 > gpurun_out/ncu_enc3.log 2>&1
python tools/ncu_summary.py gpurun_out/enc_stem7.ncu-rep > gpurun_out/enc_stem7.txt 2>&1
rm -f gpurun_out/enc_tc3.ncu-rep gpurun_out/enc_stem7.ncu-rep
ls -la gpurun_out | grep enc_
