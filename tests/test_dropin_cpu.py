"""CPU: the libsvm shim (libertext_svm.so) exports libsvm's symbols, its header restates the reference's struct layouts
exactly, and the drop-in sources name the reference lines they replace."""
import os
import subprocess
import pytest
from conftest import ROOT

PKG = os.path.join(ROOT, "scene-text-recognition_b200")
REF_INC = "/root/reference/inc"


def test_svm_shim_exports_libsvm_symbols():
    lib = os.path.join(PKG, "libertext_svm.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", PKG, "libertext_svm.so"])
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for s in ("svm_load_model", "svm_predict_probability", "svm_predict", "svm_get_nr_class", "svm_get_labels", "svm_get_svm_type", "svm_get_nr_sv",
              "svm_check_probability_model", "svm_free_model_content", "svm_free_and_destroy_model", "libsvm_version", "svm_predict_probability_batch"):
        assert (" T %s\n" % s) in syms or (" D %s\n" % s) in syms or (" B %s\n" % s) in syms, s
    assert "libertext.so" in subprocess.run(["ldd", lib], capture_output=True, text=True).stdout


def test_svm_shim_without_gpu_returns_null_like_libsvm(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except Exception:
        pass
    import ctypes
    import ertext
    ertext.load_library()
    L = ctypes.CDLL(os.path.join(PKG, "libertext_svm.so"))
    L.svm_load_model.restype = ctypes.c_void_p
    assert L.svm_load_model(ertext.svm_model_path().encode()) is None          # no CPU path: NULL, like a failed svm_load_model


LAYOUT_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include HDR
int main(void) {
	printf("%zu %zu %zu ", sizeof(struct svm_node), offsetof(struct svm_node, index), offsetof(struct svm_node, value));
	printf("%zu %zu %zu %zu %zu %zu ", sizeof(struct svm_parameter), offsetof(struct svm_parameter, gamma), offsetof(struct svm_parameter, C),
	       offsetof(struct svm_parameter, weight), offsetof(struct svm_parameter, shrinking), offsetof(struct svm_parameter, probability));
	printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(struct svm_model), offsetof(struct svm_model, nr_class), offsetof(struct svm_model, l),
	       offsetof(struct svm_model, SV), offsetof(struct svm_model, rho), offsetof(struct svm_model, sv_indices), offsetof(struct svm_model, label),
	       offsetof(struct svm_model, nSV), offsetof(struct svm_model, free_sv));
	return 0;
}
"""


def test_svm_header_layout_equals_the_reference(tmp_path):
    if not os.path.exists(os.path.join(REF_INC, "svm.h")):
        pytest.skip("/root/reference not present")
    outs = []
    for name, hdr in (("ours", os.path.join(ROOT, "include", "ertext_svm.h")), ("ref", os.path.join(REF_INC, "svm.h"))):
        src = tmp_path / ("probe_%s.c" % name)
        src.write_text(LAYOUT_PROBE.replace("HDR", '"%s"' % hdr))
        exe = tmp_path / ("probe_%s" % name)
        subprocess.check_call(["gcc", "-o", str(exe), str(src)])
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True).stdout)
    assert outs[0] == outs[1] and len(outs[0].split()) == 18


def test_dropin_defines_the_reference_signatures():
    txt = open(os.path.join(PKG, "host", "dropin", "erfilter_dropin.cpp")).read()
    for sig in ("void ERFilter::compute_channels(Mat &src, Mat &YCrcb, vector<Mat> &channels)", "ER *ERFilter::er_tree_extract(Mat input)",
                "void ERFilter::non_maximum_supression(ER *er, ERs &all, ERs &pool, Mat input)", "void ERFilter::classify(ERs &pool, ERs &strong, ERs &weak, Mat input)",
                "void ERFilter::er_track(vector<ERs> &strong, vector<ERs> &weak, ERs &all_er, vector<Mat> &channel, Mat Ycrcb)",
                "vector<double> ERFilter::make_LBP_hist(Mat input, const int N, const int normalize_size)", "Mat ERFilter::calc_LBP(Mat input, const int size)",
                "double CascadeBoost::predict(vector<double> fv)"):
        assert sig in txt, sig


def test_frame_accumulator_and_video_times_host_logic(tmp_path):
    """host/FramePipeline.hpp: FrameAccumulator = video_mode's tracked_vec over frame_count = 2 frames and the middle frame
    for er_ocr (src/utils.cpp:96-144); VideoTimes = avg_time[0..3].  Host logic only, no device call."""
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    PKG = os.path.join(ROOT, "scene-text-recognition_b200")
    exe = str(tmp_path / "acc")
    subprocess.check_call(["g++", "-std=c++11", "-O1", os.path.join(ROOT, "tests", "cpp", "accumulator_cpu.cpp"), "-o", exe,
                           "-L", PKG, "-l:libertext.so", "-Wl,-rpath," + PKG])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "push 0 ready 0 frames 1 tracked 1002 1000 middle -1"
    assert lines[1] == "push 1 ready 1 frames 2 tracked 1002 1000 1101 middle 11"       # group (10, 11): middle = the second frame
    assert lines[2] == "push 2 ready 0 frames 1 tracked middle -1"                        # a new group starts; frame 12 tracked nothing
    assert lines[3] == "push 3 ready 1 frames 2 tracked 1303 1301 1300 middle 13"
    assert lines[4] == "push 4 ready 0 frames 1 tracked 1400 middle -1"
    assert lines[5] == "times 0.005000 0.001250 0.002500 0.000625 frames 5"
    assert lines[6] == "single 1 middle 7"
