#!/bin/bash
# A/B builds of compile-time variants (results: profiles/r02_summary.md):
#   GFX_AB_NO_SPHERE    light sampling without the bounding-sphere step of sampleLightUnlessDark
#   GFX_AB_CHAIN_PICK   the three nested CDF searches instead of the flattened light pick
# Step 1 (here, no GPU):   tools/ab_flags.sh build     -> build_ab/libgfxb200_{wide,push,spheres,all}.so
# Step 2 (under gpurun):   tools/ab_flags.sh run       -> parity suite + bench line per variant (GFXB200_LIB selects the library)
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
declare -A FLAGS=( [nosphere]="-DGFX_AB_NO_SPHERE" [chainpick]="-DGFX_AB_CHAIN_PICK" [occ20]="-DGFX_RIS_MIN_BLOCKS=20" [occ24]="-DGFX_RIS_MIN_BLOCKS=24" [bounce12]="-DGFX_BOUNCE_MIN_BLOCKS=12" [bounce16]="-DGFX_BOUNCE_MIN_BLOCKS=16" )
if [ -n "${AB_ONLY:-}" ]; then for k in "${!FLAGS[@]}"; do [[ " $AB_ONLY " == *" $k "* ]] || unset "FLAGS[$k]"; done; fi
if [ "${1:-}" = "build" ]; then
    mkdir -p build_ab
    for v in "${!FLAGS[@]}"; do
        d=$(mktemp -d)
        mkdir -p "$d/gfxexp_b200" "$d/include"
        cp -r gfxexp_b200/csrc "$d/gfxexp_b200/" && cp include/*.h "$d/include/"
        (cd "$d/gfxexp_b200/csrc" && rm -f *.o && make -j8 EXTRA="${FLAGS[$v]}" > /dev/null 2>&1) || { echo "build of $v failed"; exit 1; }
        cp "$d/gfxexp_b200/libgfxb200.so" "build_ab/libgfxb200_$v.so" && rm -rf "$d"
        echo "built build_ab/libgfxb200_$v.so (${FLAGS[$v]})"
    done
elif [ "${1:-}" = "run" ]; then
    for lib in gfxexp_b200/libgfxb200.so build_ab/libgfxb200_nosphere.so build_ab/libgfxb200_chainpick.so; do
        [ -f "$lib" ] || continue
        echo "== $lib"
        GFXB200_LIB=$lib timeout -s KILL 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_host_cpp.py 2>&1 | tail -2
        GFXB200_LIB=$lib timeout -s KILL 120 python bench.py --steps 20 --no-cpu-baseline --headline-only 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['per_kernel'].items()})"
        GFXB200_LIB=$lib timeout -s KILL 120 python tools/stage_bench.py --only pathtrace 2>/dev/null | tail -1 | cut -c1-160
    done
else
    echo "usage: $0 build | run"
fi
