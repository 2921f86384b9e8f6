#!/bin/bash
# Builds a library variant with other make knobs into perturb_b200/variants/<name>.so from a
# scratch copy of csrc (so the in-tree objects and libperturb_gpu.so stay as they are):
#   scripts/build_variant.sh near8 MINB_NEAR=8
#   scripts/build_variant.sh tpl2 TPL=2
#   scripts/build_variant.sh noshort EXTRA=-DPB_SHORT_GRID=0
# then: /usr/local/graft/bin/gpurun -- 'bash scripts/sweep_variants.sh "c2 c3s c4s" 10'
# (perturb_b200/variants/ is git-ignored; the .so files travel with the gpurun snapshot).
set -e
name=$1
shift
root=$(cd "$(dirname "$0")/.." && pwd)
work=$(mktemp -d /tmp/pb_variant_XXXX)
mkdir -p "$work/perturb_b200" "$work/include" "$root/perturb_b200/variants"
cp -r "$root/perturb_b200/csrc" "$work/perturb_b200/"
cp -r "$root/include/." "$work/include/"
cd "$work/perturb_b200/csrc"
rm -f *.o
make product OUT="$root/perturb_b200/variants/$name.so" "$@" > build.log 2>&1 || { tail -20 build.log; exit 1; }
echo "$name: $(grep -c 'spill stores' sgp4_prop.ptxas.log) kernels;" \
     "near-Earth states kernel: $(grep -A2 'sgp4_prop_kernelILi0ELi0ELb0' sgp4_prop.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
rm -rf "$work"
