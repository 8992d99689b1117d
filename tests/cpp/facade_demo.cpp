// facade_demo.cpp -- exercises the C++ facade the way the reference's callers use ERFilter
// (src/utils.cpp:115-140 video_mode: compute_channels, then per channel er_tree_extract ->
// non_maximum_supression -> classify; src/utils.cpp:49 image_mode: text_detect).
// usage: facade_demo <bgr.raw> <planes6.raw> <w> <h> <strong.classifier> <weak.classifier> [OCR.model]
#include "../../scene-text-recognition_b200/host/ERFilter.hpp"
#include <cstdio>
#include <cmath>
#include <fstream>

using namespace ertx;

static std::vector<unsigned char> slurp(const char *p)
{
	std::ifstream f(p, std::ios::binary);
	return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static unsigned long long tree_hash(ER *root)
{
	unsigned long long h = 1469598103934665603ull;
	std::vector<ER *> st; st.push_back(root);
	while (!st.empty()) {
		ER *e = st.back(); st.pop_back();
		const int v[6] = {e->level, e->area, e->bound.x, e->bound.y, e->bound.width, e->bound.height};
		for (int i = 0; i < 6; i++) { h ^= (unsigned long long)(unsigned)v[i]; h *= 1099511628211ull; }
		std::vector<ER *> ch;
		for (ER *c = e->child; c; c = c->next) ch.push_back(c);
		for (int i = (int)ch.size() - 1; i >= 0; i--) st.push_back(ch[(size_t)i]);
	}
	return h;
}

int main(int argc, char **argv)
{
	if (argc < 7) { fprintf(stderr, "usage\n"); return 2; }
	const int w = atoi(argv[3]), h = atoi(argv[4]);
	std::vector<unsigned char> bgr = slurp(argv[1]), planes = slurp(argv[2]);
	if ((int)bgr.size() != w * h * 3 || (int)planes.size() != w * h * 6) { fprintf(stderr, "bad input sizes\n"); return 2; }
	try {
		ERFilter *er_filter = new ERFilter(8, 120, 900000, 2, 0.7, 0.15);                    // src/main.cpp:22
		er_filter->stc = new CascadeBoost(er_filter->device(), ERT_CASCADE_STRONG, argv[5]);  // src/main.cpp:23
		er_filter->wtc = new CascadeBoost(er_filter->device(), ERT_CASCADE_WEAK, argv[6]);    // src/main.cpp:24
		// image_mode
		Mat src(h, w, 3, bgr.data());
		ERs root; std::vector<ERs> all, pool, strong, weak;
		std::vector<double> times = er_filter->text_detect(src, root, all, pool, strong, weak);
		for (int p = 0; p < 6; p++)
			printf("TD %d pool %zu strong %zu weak %zu hash %llu\n", p, pool[(size_t)p].size(), strong[(size_t)p].size(), weak[(size_t)p].size(), tree_hash(root[(size_t)p]));
		// video_mode style: stage by stage on the channel planes
		for (int p = 0; p < 6; p++) {
			Mat ch(h, w, 1, planes.data() + (size_t)p * w * h);
			ER *r = er_filter->er_tree_extract(ch);
			ERs a, pl, s, wk;
			er_filter->non_maximum_supression(r, a, pl, ch);
			er_filter->classify(pl, s, wk, ch);
			printf("ST %d pool %zu strong %zu weak %zu hash %llu\n", p, pl.size(), s.size(), wk.size(), tree_hash(r));
			if (p == 0 && !pl.empty()) {
				std::vector<double> fv = er_filter->make_LBP_hist(ch(pl[0]->bound));
				double sum = 0; for (double v : fv) sum += v;
				printf("FV %g %g\n", sum, er_filter->stc->predict(fv));
			}
			er_filter->er_delete(r);
		}
		// er_track the way video_mode calls it after the per-channel loop (src/utils.cpp:140)
		{
			ERs tracked;
			er_filter->er_track(strong, weak, tracked, src);
			unsigned long long h2 = 1469598103934665603ull;
			for (ER *e : tracked) { const int v[5] = {e->ch, e->bound.x, e->bound.y, e->center.x, e->center.y}; for (int i = 0; i < 5; i++) { h2 ^= (unsigned long long)(unsigned)v[i]; h2 *= 1099511628211ull; } }
			printf("TR %zu hash %llu\n", tracked.size(), h2);
			// text_detect with `tracked` (one submission) must give the same list
			ERs root2, tracked2; std::vector<ERs> all2, pool2, strong2, weak2;
			std::vector<double> t2 = er_filter->text_detect(src, root2, all2, pool2, strong2, weak2, tracked2);
			unsigned long long h3 = 1469598103934665603ull;
			for (ER *e : tracked2) { const int v[5] = {e->ch, e->bound.x, e->bound.y, e->center.x, e->center.y}; for (int i = 0; i < 5; i++) { h3 ^= (unsigned long long)(unsigned)v[i]; h3 *= 1099511628211ull; } }
			printf("XTR %zu hash %llu t3 %d\n", tracked2.size(), h3, t2[3] > 0 ? 1 : 0);
			if (argc > 7) {
				OCR *ocr = new OCR(er_filter->device(), argv[7], 30, 15);                     // src/main.cpp:25
				for (size_t i = 0; i < tracked2.size() && i < 8; i++) {
					ER *e = tracked2[i];
					Mat chn(h, w, 1, planes.data() + (size_t)e->ch * w * h);
					const double result = ocr->chain_run(chn(e->bound), e->level * 8, 0.0);    // src/ER.cpp:732
					printf("OCR %c %.6f\n", (char)floor(result), result - floor(result));
				}
				delete ocr;
			}
			for (int p = 0; p < 6; p++) er_filter->er_delete(root2[(size_t)p]);
		}
		for (int p = 0; p < 6; p++) er_filter->er_delete(root[(size_t)p]);
		printf("times %d\n", (int)times.size());
		// reference error behaviour: CV_Assert on a non-8UC1 input throws
		try { er_filter->er_tree_extract(src); printf("ASSERT missing\n"); } catch (const std::exception &) { printf("ASSERT ok\n"); }
		delete er_filter->stc; delete er_filter->wtc; delete er_filter;
	} catch (const std::exception &e) { fprintf(stderr, "exception: %s\n", e.what()); return 1; }
	return 0;
}
