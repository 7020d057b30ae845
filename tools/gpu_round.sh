# One GPU-box visit: bench, per-launch profile, ncu launch list, ncu full capture of the lookup kernel, GPU tests.
# Every step has its own timeout and writes under gpurun_out/ (kept small: no whole-step --set full capture).
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 120 python tools/step_profile.py --all > gpurun_out/step_profile.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:bflow:: --launch-skip 231 -c 231 --csv --log-file gpurun_out/launches_step.csv python tools/one_step.py > gpurun_out/ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:corr_lookup --launch-skip 14 -c 1 -f -o gpurun_out/lookup_instep python tools/one_step.py > gpurun_out/ncu2.log 2>&1
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
ls -la gpurun_out
