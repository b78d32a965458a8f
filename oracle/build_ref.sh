#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the *unmodified* reference gridencoder CUDA extension
# (external/encoders/gridencoder/src/{gridencoder.cu,bindings.cpp}) straight from the
# read-only sources under /root/reference into oracle/_ref/_gridencoder_ref*.so.
# Nothing is copied into the repo. Only difference from the reference's own flags
# (external/encoders/gridencoder/setup.py:7-10): -std=c++17 instead of c++14 (torch 2.11
# headers need it) and an explicit sm_100 arch (the reference passes none).
# The .so is git-ignored but travels to the GPU box with gpurun; it is used ONLY by
# tests/ (GPU parity: our kernel vs the reference kernel, bit for bit) and by
# bench.py's optional reference-kernel comparison line.
set -euo pipefail
REF=${REF:-/root/reference/external/encoders/gridencoder/src}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF" ]; then echo "reference sources not present at $REF; skipping" >&2; exit 0; fi
PY=${PYTHON:-python}
TORCH_INC=$($PY - <<'PY'
import torch.utils.cpp_extension as c, sysconfig
print(" ".join("-I"+p for p in c.include_paths("cuda")+[sysconfig.get_paths()["include"]]))
PY
)
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
NAME=_gridencoder_ref
COMMON="-O3 -std=c++17 -DTORCH_EXTENSION_NAME=$NAME -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 $TORCH_INC"
nvcc -c "$REF/gridencoder.cu" -o "$OUT/gridencoder.o" $COMMON -Xcompiler -fPIC \
  -gencode arch=compute_100,code=sm_100 \
  -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__ \
  --expt-relaxed-constexpr
g++ -c "$REF/bindings.cpp" -o "$OUT/bindings.o" $COMMON -fPIC
g++ -shared "$OUT/gridencoder.o" "$OUT/bindings.o" -o "$OUT/$NAME$EXT" \
  -L"$TORCH_LIB" -L/usr/local/cuda/lib64 -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python -lcudart \
  -Wl,-rpath,"$TORCH_LIB"
rm -f "$OUT/gridencoder.o" "$OUT/bindings.o"
echo "built $OUT/$NAME$EXT"
