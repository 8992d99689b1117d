// grouping_demo.cpp -- the facade's er_grouping / suppressions / slope fit / er_ocr duplicate removal (host-only logic) on
// ER rows read from a text file, printed in a form tests/test_grouping_cpu.py compares with the reference's own code.
// input: first line "n overlap_sup inner_sup dedupe", then n rows: ch x y w h area cx cy color1 color2 color3
#include "../../scene-text-recognition_b200/host/ERFilter.hpp"
#include <cstdio>

using namespace ertx;

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	FILE *f = fopen(argv[1], "r");
	if (!f) return 2;
	int n, osup, isup, dedupe;
	if (fscanf(f, "%d %d %d %d", &n, &osup, &isup, &dedupe) != 4) return 2;
	std::vector<ER> E((size_t)n);
	ERs all_er;
	for (int i = 0; i < n; i++) {
		int ch, x, y, w, h, area, cx, cy;
		double c1, c2, c3;
		if (fscanf(f, "%d %d %d %d %d %d %d %d %lf %lf %lf", &ch, &x, &y, &w, &h, &area, &cx, &cy, &c1, &c2, &c3) != 11) return 2;
		E[(size_t)i].ch = ch; E[(size_t)i].bound = Rect(x, y, w, h); E[(size_t)i].area = area; E[(size_t)i].center.x = cx; E[(size_t)i].center.y = cy;
		E[(size_t)i].color1 = c1; E[(size_t)i].color2 = c2; E[(size_t)i].color3 = c3;
		all_er.push_back(&E[(size_t)i]);
	}
	fclose(f);
	std::vector<Text> text;
	er_grouping(all_er, text, osup != 0, isup != 0);
	if (dedupe) for (int i = (int)text.size() - 1; i >= 0; i--) er_ocr_remove_duplicates(text[(size_t)i]);
	printf("AFTER");
	for (ER *e : all_er) printf(" %d", (int)(e - &E[0]));
	printf("\n");
	for (int i = 0; i < n; i++) printf("B %d %d %d %d %d %d\n", E[(size_t)i].bound.x, E[(size_t)i].bound.y, E[(size_t)i].bound.width, E[(size_t)i].bound.height, E[(size_t)i].center.x, E[(size_t)i].center.y);
	for (size_t t = 0; t < text.size(); t++) {
		printf("T %.17g", text[t].slope);
		for (ER *e : text[t].ers) printf(" %d", (int)(e - &E[0]));
		printf("\n");
	}
	return 0;
}
