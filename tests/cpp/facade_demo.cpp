// facade_demo.cpp -- exercises the C++ facade the way the reference's callers use ERFilter
// (src/utils.cpp:115-140 video_mode: compute_channels, then per channel er_tree_extract ->
// non_maximum_supression -> classify; src/utils.cpp:49 image_mode: text_detect).
// usage: facade_demo <bgr.raw> <planes6.raw> <w> <h> <strong.classifier> <weak.classifier>
#include "../../scene-text-recognition_b200/host/ERFilter.hpp"
#include <cstdio>
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
		for (int p = 0; p < 6; p++) er_filter->er_delete(root[(size_t)p]);
		printf("times %d\n", (int)times.size());
		// reference error behaviour: CV_Assert on a non-8UC1 input throws
		try { er_filter->er_tree_extract(src); printf("ASSERT missing\n"); } catch (const std::exception &) { printf("ASSERT ok\n"); }
		delete er_filter->stc; delete er_filter->wtc; delete er_filter;
	} catch (const std::exception &e) { fprintf(stderr, "exception: %s\n", e.what()); return 1; }
	return 0;
}
