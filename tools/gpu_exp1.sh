mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 200 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err
BFLOW_PDL=0 timeout 200 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
{
for dbg in 0 1 2; do
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 256 --kh 1 --kw 5 --bn 128 --dbg $dbg
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 128 --kh 1 --kw 5 --bn 64 --dbg $dbg
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 128 --kh 3 --kw 3 --bn 64 --dbg $dbg
timeout 60 python tools/conv_bench.py --backend tc3 --n 5 --h 240 --w 320 --cin 64 --cout 64 --kh 3 --kw 3 --bn 64 --dbg $dbg
done
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 256 --kh 1 --kw 5 --bn 128 --trace
timeout 60 python tools/conv_bench.py --backend tc3 --cin 256 --cout 128 --kh 3 --kw 3 --bn 64 --trace
timeout 60 python tools/conv_bench.py --backend tc3 --n 5 --h 240 --w 320 --cin 64 --cout 64 --kh 3 --kw 3 --bn 64 --trace
} > gpurun_out/conv_exp1.txt 2>&1
