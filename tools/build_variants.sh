#!/bin/bash
# Developer tool: builds library variants next to the default one for tools/stage_sweep.py (`name:lib=libmgatk2_b200_<name>.so`).
#   bash tools/build_variants.sh "name:-DMACRO=1 -DOTHER=0" ...
cd "$(dirname "$0")/../mgatk2_b200/csrc" || exit 1
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static $flags \
       -o ../libmgatk2_b200_$name.so api.cu 2>&1 | grep -E "error" ; echo "built $name ($flags)"
done
