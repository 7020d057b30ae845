#!/bin/bash
# tools/mkvariant.sh NAME 'python patch code operating on variable s (conv_tc3.cu text)'   -> builds a copy of the tree under _v_NAME/
set -e
name=$1
rm -rf _v_$name && mkdir -p _v_$name
tar --exclude=.git --exclude='./_r1' --exclude='./_v_*' --exclude=gpurun_out --exclude='*.so' --exclude='*.so.sha256' --exclude=bflow_b200/build --exclude=__pycache__ --exclude=oracle/_ref -cf - . | tar -xf - -C _v_$name
cd _v_$name
python - "$2" <<'PY'
import sys
p='bflow_b200/csrc/conv_tc3.cu'
s=open(p).read()
exec(sys.argv[1])
open(p,'w').write(s)
PY
PYTHONPATH=$PWD python -m bflow_b200.build > /dev/null 2>&1 && echo "built _v_$name"
