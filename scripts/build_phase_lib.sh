#!/bin/bash
# Development: a second copy of the library with the per-phase cycle counters of the heavy numeric kernel
# (-DNSP_PHASE_TIMING), for scripts/explore_spgemm.py --phases:  NSP_LIB_PATH=build/phase/libnsparse_b200.so
set -e
cd "$(dirname "$0")/.."
make -j8 OBJ=build/phase/obj LIBDIR=build/phase EXTRA=-DNSP_PHASE_TIMING build/phase/libnsparse_b200.so
