#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt
tail -15 gpurun_out/r2b_pytest.txt
timeout 300 python tools/ref_tf32_on_fixture.py d_480x640_i12 m_384x512_i12 d_128_i4_bn > gpurun_out/r2b_ref_tf32.txt 2>&1
cat gpurun_out/r2b_ref_tf32.txt | grep reference
{
for dbg in 0 1; do
  python tools/conv_bench.py --cin 256 --cout 192 --kh 3 --kw 3 --bn 64 --backend tc3 --dbg $dbg --trace
  python tools/conv_bench.py --cin 256 --cout 256 --kh 1 --kw 5 --bn 128 --backend tc3 --dbg $dbg --trace
done
} > gpurun_out/r2b_convbench.txt 2>&1
grep "TF/s" gpurun_out/r2b_convbench.txt
