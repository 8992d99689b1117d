// video_demo.cpp -- video_mode's loop (src/utils.cpp:101-216) on compressed input: JPEG frames go in one by one
// (FramePipeline::push_jpeg, decoded on the device), per-frame regions come out in order, FrameAccumulator collects the
// tracked regions of frame_count = 2 consecutive frames (tracked_vec, src/utils.cpp:143-144) and VideoTimes keeps avg_time[7].
// usage: video_demo <frames.jpgs> <sizes.txt> <w> <h> <n_total> <frames_per_batch> <depth> <strong> <weak>
#include "../../scene-text-recognition_b200/host/FramePipeline.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

int main(int argc, char **argv)
{
	if (argc < 10) { fprintf(stderr, "usage\n"); return 2; }
	const int w = atoi(argv[3]), h = atoi(argv[4]), total = atoi(argv[5]), fpb = atoi(argv[6]), depth = atoi(argv[7]);
	std::ifstream f(argv[1], std::ios::binary);
	std::vector<unsigned char> blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	std::vector<size_t> off(1, 0);
	{ std::ifstream s(argv[2]); size_t n; while (s >> n) off.push_back(off.back() + n); }
	const int nd = (int)off.size() - 1;
	if (nd < 1 || off.back() != blob.size()) { fprintf(stderr, "bad input\n"); return 2; }
	try {
		ert_params prm = {8, 120, 900000, 2, 0.7, 0.15};
		ertx::FramePipeline pipe(prm, 0, w, h, argv[8], argv[9], fpb, depth);
		ertx::FrameAccumulator acc(2);
		ertx::VideoTimes times;
		ertx::FrameRegions out;
		long long groups = 0;
		auto consume = [&](ertx::FrameRegions &r) {
			times.add(r);
			if (!acc.push(std::move(r))) return;
			unsigned long long hs = 1469598103934665603ull;
			for (const ert_tracked &c : acc.tracked()) {
				const int v[5] = {c.plane, c.x, c.y, c.center_x, c.center_y};
				for (int k = 0; k < 5; k++) { hs ^= (unsigned long long)(unsigned)v[k]; hs *= 1099511628211ull; }
			}
			// here video_mode calls er_grouping(tracked_vec, ...) and er_ocr(tracked_vec, channel_vec, ...) -- the reference's own host code
			printf("G %lld middle %lld tracked %zu hash %llu\n", groups, acc.middle_frame_index(), acc.tracked().size(), hs);
			groups++;
		};
		const auto t0 = std::chrono::high_resolution_clock::now();
		for (int i = 0; i < total; i++) {
			const int k = i % nd;
			pipe.push_jpeg(blob.data() + off[(size_t)k], off[(size_t)k + 1] - off[(size_t)k]);
			while (pipe.pop(out, false)) consume(out);
		}
		pipe.flush();
		while (pipe.pop(out)) consume(out);
		times.t[6] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		printf("TIMES extract %.6f nms %.6f classify %.6f track %.6f grouping %.6f ocr %.6f total %.6f frames %lld\n", times.t[0], times.t[1], times.t[2], times.t[3],
		       times.t[4], times.t[5], times.t[6], times.frames);
	} catch (const std::exception &e) { fprintf(stderr, "exception: %s\n", e.what()); return 1; }
	return 0;
}
