#!/usr/bin/env bash
# Compiles the REFERENCE's own sources, untouched, from where they lie under /root/reference into
# oracle/_ref/*.so (git-ignored, not gpurun-ignored).  No reference source is copied into the repo:
# whole files are compiled in place; for tf_interpolate.cpp only the TF-free line range 60-153
# (threenn_cpu / threeinterpolate_cpu / threeinterpolate_grad_cpu) is piped straight into g++.
# The reference's own build system (compile_op.sh, needs TensorFlow) is NOT run.
set -euo pipefail
REF="${PASNL_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/tf_ops" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# (1) GPU oracle for FPS / gather / ball query / group / selection sort: reference .cu files as they are.
#     The reference builds with -O2 and no -arch (tf_ops/compile_ops.sh:10,13); we only add the sm_100a target.
"$NVCC" -O2 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
    -o "$OUT/ref_sampling.so" "$REF/tf_ops/sampling/tf_sampling_g.cu"
"$NVCC" -O2 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
    -o "$OUT/ref_grouping.so" "$REF/tf_ops/grouping/tf_grouping_g.cu"
# (2) CPU oracle for kNN: knn_.cxx + vendored nanoflann 1.2.3 (utils/nearest_neighbors/setup.py uses -fopenmp, -std=c++11).
g++ -O2 -std=c++11 -fopenmp -fPIC -shared -I"$REF/utils/nearest_neighbors" \
    -o "$OUT/ref_knn.so" "$REF/utils/nearest_neighbors/knn_.cxx"
# (3) CPU oracle for three_nn / three_interpolate(+grad): tf_interpolate.cpp:60-153, TF-free plain C++.
{
  echo '#include <cstdio>'
  echo '#include <cstring>'
  echo '#include <cmath>'
  echo 'extern "C" {'
  sed -n '60,153p' "$REF/tf_ops/3d_interpolation/tf_interpolate.cpp"
  echo '}'
} | g++ -O2 -fPIC -shared -x c++ - -o "$OUT/ref_interp.so"
# (4) CPU oracle for grid subsampling: the reference's grid_subsampling.cpp + cloud.cpp as they are, plus a 30-line
#     extern "C" shim of ours (oracle/ref_shims/) in place of the CPython wrapper (wrapper.cpp needs numpy headers).
CW="$REF/utils/cpp_wrappers"
g++ -O2 -std=c++11 -fPIC -shared -I"$CW/cpp_subsampling" -o "$OUT/ref_gridsub.so" \
    "$HERE/ref_shims/grid_subsampling_shim.cpp" "$CW/cpp_subsampling/grid_subsampling/grid_subsampling.cpp" \
    "$CW/cpp_utils/cloud/cloud.cpp"
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
