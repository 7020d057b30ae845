mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
{
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 128 --kh 3 --kw 3 --bn 64 --trace
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 256 --kh 3 --kw 3 --bn 128
} > gpurun_out/mma_exp.txt 2>&1
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline22.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench22.json 2> gpurun_out/bench22.err
