#!/bin/bash
# A/B builds of libren_b200.so that differ only in csrc/ren_tc.cu compile-time switches (the other objects are the
# in-tree build's).  Usage: tools/build_variants.sh  ->  robust-e-nerf_b200/ren_b200/lib/variants/libren_b200_<tag>.so
#   bias    : REN_BIAS_MMA=1 (hidden-layer biases as one more K step of the GEMM)
#   nosplit : REN_WIDE_SPLIT=0 (one N = 256 MMA instead of an interleaved N = 128 pair)
#   ahead2  : FLAT_AHEAD=2 (forward flat epilogue: softplus two groups ahead of the split / stores)
#   nohint  : REN_MBAR_HINT_NS=0 (mbarrier waits spin instead of suspending)
#   noelect : REN_ELECT=0 (single-thread roles on lane 0 instead of an elect.sync lane)
#   generic : REN_FLAT_FWD=0 REN_FLAT_BWD=0 (generic epilogue loop only)
#   old     : csrc/ren_tc.cu of commit e330866 (before the flat epilogues)
# Select one at run time with REN_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../robust-e-nerf_b200"
python -c "import sys; sys.path.insert(0, '..'); import __graft_entry__ as g; g.build()"
rm -rf ren_b200/lib/variants build/variants
mkdir -p ren_b200/lib/variants build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OTHERS=$(ls build/*.o | grep -v "build/ren_tc")
TC="ren_tc ren_tc_wgrad ren_tc_raygrad"
build() {  # tag, defines...: the three tcgen05 sources with the defines, everything else from the in-tree build
  tag=$1; shift
  objs=""
  for f in $TC; do
    nvcc $FLAGS "$@" -c csrc/$f.cu -o build/variants/${f}_$tag.o
    objs="$objs build/variants/${f}_$tag.o"
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ren_b200/lib/variants/libren_b200_$tag.so \
      $objs $OTHERS -lcuda
}
build bias -DREN_BIAS_MMA=1 &
build nosplit -DREN_WIDE_SPLIT=0 &
build ahead2 -DFLAT_AHEAD=2 &
build nohint -DREN_MBAR_HINT_NS=0 &
build noelect -DREN_ELECT=0 &
build generic -DREN_FLAT_FWD=0 -DREN_FLAT_BWD=0 &
# old: the chain kernel as committed before the flat epilogues (e330866), for the before / after timing
git show e330866:robust-e-nerf_b200/csrc/ren_tc.cu > build/variants/ren_tc_old.cu
(nvcc $FLAGS -DREN_ELECT=0 -Icsrc -I../include -c build/variants/ren_tc_old.cu -o build/variants/ren_tc_old.o &&
 nvcc $FLAGS -DREN_ELECT=0 -c csrc/ren_tc_wgrad.cu -o build/variants/ren_tc_wgrad_old.o &&
 nvcc $FLAGS -DREN_ELECT=0 -c csrc/ren_tc_raygrad.cu -o build/variants/ren_tc_raygrad_old.o &&
 nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ren_b200/lib/variants/libren_b200_old.so \
     build/variants/ren_tc_old.o build/variants/ren_tc_wgrad_old.o build/variants/ren_tc_raygrad_old.o $OTHERS -lcuda) &
wait
ls -la ren_b200/lib/variants
