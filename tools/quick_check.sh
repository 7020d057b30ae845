# GPU tests + a short bench + the in-graph timeline of the committed state (about one minute on the box)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 200 python bench.py --no-sweep > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline_quick.txt 2>&1
