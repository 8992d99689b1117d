// dropin_main.cpp -- TEST HARNESS for oracle/build_dropin_test.sh (see there): drives the reference's OWN
// ERFilter::text_detect (src/ER.cpp:33-111), the per-frame block of video_mode (src/utils.cpp:113-141) and OCR::chain_run
// (src/OCR.cpp:67-140), all compiled unmodified, with this repository's drop-in definitions underneath, and prints what
// they produce so that tests/test_gpu_dropin.py can compare it with the reference's own results (tests/golden).
//
//   dropin_demo <frames.bin> <strong.classifier> <weak.classifier> [<OCR.model> <ocr_rows.txt>]
// frames.bin = int32 n, h, w followed by n * h * w * 3 BGR bytes.  ocr_rows.txt = lines "frame plane x y w h slope".
#include "ER.h"
#include <cstdio>
#include <cstdlib>

static void print_regions(const char *tag, int frame, vector<ERs> &v)
{
	for (size_t ch = 0; ch < v.size(); ch++)
		for (size_t i = 0; i < v[ch].size(); i++) {
			ER *e = v[ch][i];
			printf("%s %d %d %d %d %d %d %d\n", tag, frame, (int)ch, e->bound.x, e->bound.y, e->bound.width, e->bound.height, e->area);
		}
}

static void print_tracked(const char *tag, int frame, ERs &t)
{
	for (size_t i = 0; i < t.size(); i++) {
		ER *e = t[i];
		printf("%s %d %d %d %d %d %d %d %d %d %.17g %.17g %.17g\n", tag, frame, e->ch, e->bound.x, e->bound.y, e->bound.width, e->bound.height, e->area,
		       e->center.x, e->center.y, e->color1, e->color2, e->color3);
	}
}

// the per-frame block of video_mode, verbatim (src/utils.cpp:113-141), inside the variables it expects
static void video_frame(ERFilter *er_filter, Mat frame, int frame_index)
{
	const int frame_count = 2;
	const int n = 1;
	vector<Mat> channel_vec;
#include "ref_video_frame.inc"
	print_regions("vstrong", frame_index, strong);
	print_regions("vweak", frame_index, weak);
	print_tracked("vtracked", frame_index, tracked);
	printf("vchannel_vec %d %d\n", frame_index, (int)channel_vec.size());
	for (size_t i = 0; i < root.size(); i++) er_filter->er_delete(root[i]);
}

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: dropin_demo frames.bin strong.classifier weak.classifier [OCR.model ocr_rows.txt]\n"); return 2; }
	FILE *f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	int hdr[3];
	if (fread(hdr, sizeof(int), 3, f) != 3) return 2;
	const int n = hdr[0], h = hdr[1], w = hdr[2];
	std::vector<unsigned char> data((size_t)n * h * w * 3);
	if (fread(data.data(), 1, data.size(), f) != data.size()) return 2;
	fclose(f);

	// exactly what src/main.cpp:22-24 does
	ERFilter *er_filter = new ERFilter(8, 120, 900000, 2, 0.7, 0.15);
	er_filter->stc = new CascadeBoost(argv[2]);
	er_filter->wtc = new CascadeBoost(argv[3]);
	printf("num_iter %d %d\n", er_filter->stc->get_num_iter(), er_filter->wtc->get_num_iter());

	try {
		for (int i = 0; i < n; i++) {
			Mat src(h, w, CV_8UC3, data.data() + (size_t)i * h * w * 3, (size_t)w * 3);
			ERs root, tracked;
			vector<ERs> all, pool, strong, weak;
			vector<Text> text;
			vector<double> times = er_filter->text_detect(src, root, all, pool, strong, weak, tracked, text);   // the reference's code
			print_regions("pool", i, pool);
			print_regions("strong", i, strong);
			print_regions("weak", i, weak);
			print_tracked("tracked", i, tracked);
			for (size_t t = 0; t < text.size(); t++) {
				printf("text %d %d %.17g", i, (int)text[t].ers.size(), text[t].slope);
				for (size_t k = 0; k < text[t].ers.size(); k++) printf(" %d:%d:%d", text[t].ers[k]->ch, text[t].ers[k]->bound.x, text[t].ers[k]->bound.y);
				printf("\n");
			}
			printf("times %d %zu %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n", i, times.size(), times[0], times[1], times[2], times[3], times[4], times[5], times[6]);
			for (size_t k = 0; k < root.size(); k++) er_filter->er_delete(root[k]);
			video_frame(er_filter, src, i);
		}
		// make_LBP_hist / calc_LBP / CascadeBoost::predict through the reference's own signatures
		{
			Mat src(h, w, CV_8UC3, data.data(), (size_t)w * 3);
			Mat ycc;
			vector<Mat> ch;
			er_filter->compute_channels(src, ycc, ch);
			Mat roi = ch[0](Rect(10, 20, 57, 41));
			vector<double> fv = er_filter->make_LBP_hist(roi);
			Mat lbp = er_filter->calc_LBP(roi);
			double sum = 0; for (size_t k = 0; k < fv.size(); k++) sum += fv[k];
			int hist2[1024] = {0};
			for (int y = 0; y < 24; y++) for (int x = 0; x < 24; x++) hist2[(y / 12) * 512 + (x / 12) * 256 + lbp.ptr(y)[x]]++;
			int diff = 0; for (int k = 0; k < 1024; k++) diff += (hist2[k] != (int)fv[k]);
			printf("lbp %d %.1f %d\n", (int)fv.size(), sum, diff);
			printf("predict %.17g %.17g\n", er_filter->stc->predict(fv), er_filter->wtc->predict(fv));
		}
		if (argc >= 6) {
			// OCR::OCR and OCR::chain_run are the reference's code; svm_load_model / svm_predict_probability are libertext_svm.so's
			er_filter->ocr = new OCR(argv[4], 30, 15);
			FILE *g = fopen(argv[5], "r");
			if (!g) { perror(argv[5]); return 2; }
			int fr, pl, x, y, ww, hh; double slope;
			int last = -1;
			Mat ycc; vector<Mat> ch;
			while (fscanf(g, "%d %d %d %d %d %d %lf", &fr, &pl, &x, &y, &ww, &hh, &slope) == 7) {
				if (fr != last) { Mat src(h, w, CV_8UC3, data.data() + (size_t)fr * h * w * 3, (size_t)w * 3); er_filter->compute_channels(src, ycc, ch); last = fr; }
				const double v = er_filter->ocr->chain_run(ch[(size_t)pl](Rect(x, y, ww, hh)), 0, slope);
				printf("ocr %d %d %d %d %d %d %.17g\n", fr, pl, x, y, ww, hh, v);
			}
			fclose(g);
		}
	} catch (const std::exception &ex) {
		fprintf(stderr, "dropin_demo: %s\n", ex.what());
		return 1;
	}
	return 0;
}
