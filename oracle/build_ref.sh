#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libbwaref.so = the reference's own
# CPU implementation of the hot path, compiled from the sources where they lie under
# /root/reference (nothing is copied into this repo), plus our thin driver
# oracle/ref_driver.cpp.  Recipe verified in SURVEY.md Appendix A:
#   -fcommon          rle_auxtab is defined in a header (bwa/rle.h:33)
#   -fgnu89-inline    non-static inline ks_getc/ks_getuntil/get_rlen/mem_infer_dir
#                     (bwa/kseq.h:71,144  bwa/bwamem.c:810  bwa/bwamem_pair.c:23)
#   USE_STD_BWA undefined  keeps smem_aux_init/test_and_merge exported
#   USE_HTSLIB  undefined  text-SAM path of src/bwa_wrapper.cpp:579-590
#   bwamem.c only: -Dksw_extend2=bwaref_ksw_hook observes the extension jobs of
#                     mem_chain2aln (bwa/bwamem.c:740,768) without touching the source.
set -euo pipefail
REF=${BWAFLOW_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
if [ ! -d "$REF/bwa" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt $OUT (if any)" >&2
  exit 0
fi
if [ -f "$OUT/libbwaref.so" ] && [ "$OUT/libbwaref.so" -nt "$HERE/ref_driver.cpp" ] && [ "$OUT/libbwaref.so" -nt "$HERE/build_ref.sh" ]; then
  exit 0
fi
mkdir -p "$OBJ"
CFLAGS="-O2 -g0 -fPIC -fcommon -fgnu89-inline -w -DHAVE_PTHREAD -DUSE_MALLOC_WRAPPERS -I$HERE/stub"
pids=()
for f in utils kthread kstring ksw bwt bntseq bwa bwamem bwamem_pair bwamem_extra malloc_wrap \
         QSufSort bwt_gen rope rle is bwtindex kopen; do
  extra=""
  [ "$f" = bwamem ] && extra="-Dksw_extend2=bwaref_ksw_hook"
  [ "$f" = bwamem_pair ] && extra="-Dksw_align2=bwaref_sw_hook"
  gcc -c $CFLAGS $extra "$REF/bwa/$f.c" -o "$OBJ/$f.o" &
  pids+=($!)
done
g++ -c -O2 -g0 -fPIC -std=c++11 -fpermissive -w -I"$HERE/stub" -I"$REF" -I"$REF/src" \
    "$REF/src/bwa_wrapper.cpp" -o "$OBJ/bwa_wrapper.o" &
pids+=($!)
g++ -c -O2 -g0 -fPIC -std=c++11 -fpermissive -w -I"$HERE/stub" -I"$REF" -I"$REF/src" \
    "$HERE/ref_driver.cpp" -o "$OBJ/ref_driver.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$OUT/libbwaref.so" "$OBJ"/*.o -lz -lpthread -lm
echo "[build_ref] built $OUT/libbwaref.so" >&2
