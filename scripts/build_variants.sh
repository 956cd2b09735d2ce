#!/bin/bash
# Builds tuning variants of libroboplay.so into robopianist_b200/variants/ (not committed; they travel with gpurun).
# usage: build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
set -e
cd "$(dirname "$0")/../robopianist_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -shared -o ../variants/lib_$name.so rp_host.cu rp_kernels.cu rp_task_kernels.cu 2> ../variants/build_$name.log
  echo "$name: $(grep -A1 'rp_step_kernelILi2' ../variants/build_$name.log | grep -o 'Used [0-9]* registers') $(grep -c . ../variants/build_$name.log >/dev/null; cuobjdump -sass ../variants/lib_$name.so | grep -E 'Function :|^\s+/\*[0-9a-f]{4,}\*/' | awk '/Function/{name=$3; next}{c[name]++}END{for(k in c) if (k ~ /ILi2/) print c[k], "sass"}')"
done
