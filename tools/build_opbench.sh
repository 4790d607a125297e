#!/bin/bash
# Builds tools/opbench (native timing / A-B harness over the C ABI) next to the library it links.
set -e
cd "$(dirname "$0")/.."
python -m lgteun_b200.build >/dev/null
nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/opbench tools/opbench.cu \
     -L lgteun_b200 -l:_lgteun_cuda.so -Xlinker -rpath -Xlinker '$ORIGIN/../lgteun_b200'
echo tools/opbench
