#!/usr/bin/env bash
# End-to-end throughput of the C++ streaming runtime (host/FramePipeline.hpp) on the bench workload, no Python in the loop:
# 8 distinct synthetic 1080p S-text frames pushed round-robin, results popped per frame.  usage: tools/stream_bench.sh [total] [fpb] [depth]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/scene-text-recognition_b200"
TOTAL="${1:-800}"; FPB="${2:-8}"; DEPTH="${3:-5}"
TMP="$(mktemp -d)"
g++ -std=c++11 -O2 "$ROOT/tests/cpp/stream_demo.cpp" -o "$TMP/stream_demo" -L "$PKG" -l:libertext.so -Wl,-rpath,"$PKG"
python - "$TMP/frames.raw" <<'PY'
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(sys.argv[0]))) if False else None
sys.path.insert(0, os.path.join(os.environ.get("ERT_ROOT", "."), "scene-text-recognition_b200"))
from ertext import synth
synth.s_text_batch(1234, 8, 1920, 1080).tofile(sys.argv[1])
PY
for MODE in 1 2; do
	echo -n "mode $MODE (1 = frames copied into staging by the caller thread, 2 = written in place): "
	"$TMP/stream_demo" "$TMP/frames.raw" 8 1920 1080 "$TOTAL" "$FPB" "$DEPTH" "$ROOT/assets/classifier/strong.classifier" "$ROOT/assets/classifier/weak.classifier" $MODE | tail -1
done
rm -rf "$TMP"
