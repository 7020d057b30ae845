mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "slab_mode" 2>&1 | tail -30 > gpurun_out/pytest_slabmode.txt
if grep -q passed gpurun_out/pytest_slabmode.txt && ! grep -q failed gpurun_out/pytest_slabmode.txt; then
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline17.txt 2>&1
timeout 200 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err
BFLOW_TC3_SLAB=0 timeout 200 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/bench17_noslab.json 2> gpurun_out/bench17_noslab.err
fi
