#!/usr/bin/env bash
# Build recipe for oracle/_ref: the UNMODIFIED reference CUDA extension (svox_t/csrc) compiled for
# sm_100a, used by tests/ and bench.py only as a checker ("the kernel to beat on the same box").
#
# TEST INFRASTRUCTURE ONLY. Nothing under svox_t_b200/ may import or link anything produced here.
#
# The reference sources stay where they lie (/root/reference). They do not compile against
# torch >= 2.x because 23 call sites pass `x.type()` to AT_DISPATCH_FLOATING_TYPES; the build therefore
# works on a scratch copy under /tmp on which one mechanical sed (`.type()` -> `.scalar_type()` inside
# the dispatch macro) is applied. Only the resulting .so lands in oracle/_ref/ (git-ignored, shipped
# to the GPU box by gpurun). No reference source is copied into this repository.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
SCRATCH=${SCRATCH:-/tmp/svox_t_ref_build}
if [ ! -d "$REF/svox_t/csrc" ]; then
  echo "build_ref: $REF/svox_t/csrc not present (GPU box?) - using prebuilt oracle/_ref if any"; exit 0
fi
rm -rf "$SCRATCH" && mkdir -p "$SCRATCH/build" "$OUT"
cp -r "$REF/svox_t/csrc" "$SCRATCH/csrc"
chmod -R u+w "$SCRATCH"
sed -i -E 's/(AT_DISPATCH_FLOATING_TYPES\([A-Za-z_.]+)\.type\(\)/\1.scalar_type()/' \
    "$SCRATCH"/csrc/*.cu "$SCRATCH"/csrc/*.cpp
cd "$SCRATCH"
TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=${MAX_JOBS:-8} python - <<PY
from torch.utils.cpp_extension import load
import glob, os, shutil
src = sorted(glob.glob("$SCRATCH/csrc/*.cu") + glob.glob("$SCRATCH/csrc/*.cpp"))
load(name="svox_t_ref_csrc", sources=src, extra_include_paths=["$SCRATCH/csrc/include"],
     build_directory="$SCRATCH/build", verbose=False, is_python_module=False)
so = "$SCRATCH/build/svox_t_ref_csrc.so"
assert os.path.exists(so), so
shutil.copy(so, "$OUT/svox_t_ref_csrc.so")
print("built", "$OUT/svox_t_ref_csrc.so", os.path.getsize(so) >> 20, "MiB")
PY
