#!/bin/sh
# Regenerates tests/golden/varstore_libtorch.ot (needs the torch wheel's C++ headers / libs).
set -e
TORCH=$(python -c "import torch,os; print(os.path.dirname(torch.__file__))")
g++ -O1 -std=c++17 tests/golden/make_varstore_fixture.cpp -I$TORCH/include -I$TORCH/include/torch/csrc/api/include \
    -L$TORCH/lib -ltorch -ltorch_cpu -lc10 -Wl,-rpath,$TORCH/lib -o /tmp/make_varstore_fixture
/tmp/make_varstore_fixture tests/golden/varstore_libtorch.ot
