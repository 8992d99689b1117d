#!/usr/bin/env bash
# oracle/build_ref.sh -- build the REFERENCE-BACKED oracle (test infrastructure only).
#
# Compiles the reference's own hot-path code from where it lies under /root/reference:
#   * src/adaboost.cpp, src/svm.cpp            -- unmodified, whole files
#   * src/ER.cpp, src/OCR.cpp                  -- the hot-path function bodies, extracted by
#     line range (they sit in translation units that also hold GUI / training code which needs
#     the full OpenCV SDK; the reference's CMake build fails here at find_package(OpenCV)).
# against the reference's own headers (inc/*.h, unmodified) and oracle/cvshim (a minimal
# stand-in for <opencv2/opencv.hpp>).  Outputs go ONLY to oracle/_ref/ (git-ignored); no
# reference source is copied into the repository.
#
# Line ranges (reference commit 1025d09):
#   ER.cpp   6-10    ER::ER
#            14-30   ERFilter::ERFilter, set_thresh_step, set_min_area
#            131-233 er_accumulate, er_merge, er_delete
#            240-413 er_tree_extract, process_stack
#            416-528 non_maximum_supression, classify
#            789-845 make_LBP_hist, calc_LBP
#   OCR.cpp  394-430 OCR::ARAN
# and for the rows that follow the detect path (checked by tests/test_*next*.py):
#   ER.cpp   532-609   ERFilter::er_track
#            1391-1437 calc_color
#            612-692   ERFilter::er_grouping
#            893-964   ERFilter::inner_suppression, overlap_suppression
#            1361-1389 fitline_avgslope
#            702-724   the duplicate-removal loop body of ERFilter::er_ocr (emitted as an include fragment, ref_er_ocr_dedupe.inc,
#                      which ref_capi_next.cpp wraps in the loop header of src/ER.cpp:700-701)
#   OCR.cpp  4-15      enum category, table[], cat[]
#            18-21     OCR::OCR(model, img_L, feature_L)
#            67-140    OCR::chain_run
#            144-250   OCR::extract_feature
#            254-360   OCR::rotate_mat
#            602-622   OCR::chain_code_direction
set -euo pipefail
REF="${ERT_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/ER.cpp" ]; then
	echo "build_ref.sh: $REF not present (GPU box?) -- keeping prebuilt $OUT/libref_oracle.so if any" >&2
	exit 0
fi
mkdir -p "$OUT"
{
	echo '#include "ER.h"'
	sed -n '6,10p;14,30p;131,233p;240,413p;416,528p;532,609p;612,692p;789,845p;893,964p;1361,1389p;1391,1437p' "$REF/src/ER.cpp"
	sed -n '4,15p;18,21p;67,140p;144,250p;254,360p;394,430p;602,622p' "$REF/src/OCR.cpp"
} > "$OUT/ref_hotpath.cpp"
sed -n '702,724p' "$REF/src/ER.cpp" > "$OUT/ref_er_ocr_dedupe.inc"
g++ -std=c++11 -O2 -fopenmp -fPIC -shared -w -I "$OUT" \
	-I "$HERE/cvshim" -I "$REF/inc" \
	"$OUT/ref_hotpath.cpp" "$REF/src/adaboost.cpp" "$REF/src/svm.cpp" "$HERE/ref_capi.cpp" "$HERE/ref_capi_next.cpp" \
	-o "$OUT/libref_oracle.so"
echo "built $OUT/libref_oracle.so"
