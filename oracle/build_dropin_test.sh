#!/usr/bin/env bash
# oracle/build_dropin_test.sh -- TEST INFRASTRUCTURE: the drop-in proof for SURVEY 8(b).
#
# Builds oracle/_ref/dropin_demo = the REFERENCE's own callers, extracted by line range from where they lie under
# /root/reference and compiled UNMODIFIED against the reference's own headers (inc/ER.h, inc/adaboost.h, inc/OCR.h,
# inc/svm.h), linked with this repository's drop-in translation unit (host/dropin/erfilter_dropin.cpp) and libraries
# (libertext.so, libertext_svm.so) INSTEAD of the reference's hot-path definitions:
#   ER.cpp    6-10      ER::ER
#             14-30     ERFilter::ERFilter, set_thresh_step, set_min_area
#             33-111    ERFilter::text_detect                  <- the caller under test (compiled with DO_OCR off: it then
#                                                                 ends with er_grouping; the word graph / spell check are out of scope)
#             194-233   ERFilter::er_delete
#             612-692   ERFilter::er_grouping
#             893-964   inner_suppression, overlap_suppression
#             1361-1389 fitline_avgslope
#   utils.cpp 113-141   the per-frame block of video_mode (compute_channels, the omp loop over the three stage functions,
#                       er_track), emitted as an include fragment that tests/cpp/dropin_main.cpp wraps in a function
#   adaboost.cpp        whole file minus lines 507-542 (CascadeBoost::predict, replaced by the drop-in)
#   OCR.cpp   4-15, 18-21, 67-140, 144-250, 254-360, 394-430, 602-622   OCR::OCR, chain_run, extract_feature, rotate_mat, ARAN,
#                       chain_code_direction -- unmodified; their svm_load_model / svm_predict_probability calls
#                       (src/OCR.cpp:20, 92) bind to libertext_svm.so, src/svm.cpp is NOT compiled
# OpenCV is the oracle's stand-in (oracle/cvshim).  Outputs only under oracle/_ref/ (git-ignored, travels to the GPU box).
set -euo pipefail
REF="${ERT_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
PKG="$ROOT/scene-text-recognition_b200"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/ER.cpp" ]; then
	echo "build_dropin_test.sh: $REF not present (GPU box?) -- keeping prebuilt $OUT/dropin_demo if any" >&2
	exit 0
fi
mkdir -p "$OUT/dropin"
{
	echo '#include "ER.h"'
	echo '#undef DO_OCR'
	sed -n '6,10p;14,30p;33,111p;194,233p;612,692p;893,964p;1361,1389p' "$REF/src/ER.cpp"
	sed -n '4,15p;18,21p;67,140p;144,250p;254,360p;394,430p;602,622p' "$REF/src/OCR.cpp"
} > "$OUT/dropin/ref_callers.cpp"
sed -n '113,141p' "$REF/src/utils.cpp" > "$OUT/dropin/ref_video_frame.inc"
sed '507,542d' "$REF/src/adaboost.cpp" > "$OUT/dropin/ref_adaboost_without_predict.cpp"
g++ -std=c++11 -O2 -fopenmp -w -I "$OUT/dropin" -I "$HERE/cvshim" -I "$REF/inc" \
	"$OUT/dropin/ref_callers.cpp" "$OUT/dropin/ref_adaboost_without_predict.cpp" \
	"$PKG/host/dropin/erfilter_dropin.cpp" "$ROOT/tests/cpp/dropin_main.cpp" \
	-L "$PKG" -lertext -lertext_svm -Wl,-rpath,"\$ORIGIN/../../scene-text-recognition_b200" \
	-o "$OUT/dropin_demo"
echo "built $OUT/dropin_demo"
