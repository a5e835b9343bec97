#!/bin/bash
# The reference is not a pip-installable package (no setup.py / pyproject at its root), so the base contract's
# `pip install --target baseline/_ref /root/reference` does not apply; this is its equivalent: the reference's Python package
# copied UNMODIFIED into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box, where /root/reference does
# not exist) so that `bench.py --impl reference` can time the real reference on the box's host cores through oracle/ref_loader.py.
# Prebuilt binaries and build directories of the reference's CUDA op are left out.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PDB_REFERENCE:-/root/reference}"
[ -d "$REF/part_distillation" ] || { echo "reference checkout not found at $REF" >&2; exit 3; }
mkdir -p "$HERE/_ref"
rm -rf "$HERE/_ref/part_distillation"
(cd "$REF" && find part_distillation -name '*.py' -not -path '*/ops/build/*' -not -path '*/ops/dist/*' -print0 | tar --null -T - -cf - | tar -xf - -C "$HERE/_ref")
echo "installed $(find "$HERE/_ref/part_distillation" -name '*.py' | wc -l) files into $HERE/_ref/part_distillation"
