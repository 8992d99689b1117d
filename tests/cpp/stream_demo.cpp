// stream_demo.cpp -- drives host/FramePipeline.hpp the way video_mode drives the reference (src/utils.cpp:101-216):
// frames in one by one, per-frame regions out in order.  Prints one line per frame (counts + a hash of the tracked
// list) for the pytest to compare with the Python binding, then a throughput line.
// usage: stream_demo <frames.raw> <n_distinct> <w> <h> <n_total> <frames_per_batch> <depth> <strong> <weak> [bench: 1 = copy frames in, 2 = frames written in place]
#include "../../scene-text-recognition_b200/host/FramePipeline.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

int main(int argc, char **argv)
{
	if (argc < 10) { fprintf(stderr, "usage\n"); return 2; }
	const int nd = atoi(argv[2]), w = atoi(argv[3]), h = atoi(argv[4]), total = atoi(argv[5]), fpb = atoi(argv[6]), depth = atoi(argv[7]);
	std::ifstream f(argv[1], std::ios::binary);
	std::vector<unsigned char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	const size_t fb = (size_t)w * h * 3;
	if (raw.size() != fb * (size_t)nd) { fprintf(stderr, "bad input size\n"); return 2; }
	try {
		ert_params prm = {8, 120, 900000, 2, 0.7, 0.15};
		ertx::FramePipeline pipe(prm, 0, w, h, argv[8], argv[9], fpb, depth);
		ertx::FrameRegions out;
		long long popped = 0;
		auto report = [&](const ertx::FrameRegions &r) {
			size_t np = 0, ns = 0, nw = 0;
			for (int ch = 0; ch < 6; ch++) {
				np += r.pool[ch].size();
				for (int32_t l : r.label[ch]) { ns += l == ERT_LABEL_STRONG; nw += l == ERT_LABEL_WEAK; }
			}
			unsigned long long hs = 1469598103934665603ull;
			for (int32_t i : r.tracked) {
				const ert_tracked &c = r.cand[(size_t)i];
				const int v[5] = {c.plane, c.x, c.y, c.center_x, c.center_y};
				for (int k = 0; k < 5; k++) { hs ^= (unsigned long long)(unsigned)v[k]; hs *= 1099511628211ull; }
			}
			printf("F %lld pool %zu strong %zu weak %zu tracked %zu hash %llu\n", r.frame_index, np, ns, nw, r.tracked.size(), hs);
		};
		const int bench = argc > 10 ? atoi(argv[10]) : 0;
		if (bench) {
			// untimed warm-up: every slot allocates its device workspace on first use; in mode 2 it also leaves frame
			// (i % nd) in every staging position, so the timed loop can commit frames "written in place" by a producer
			if ((fpb * depth) % nd != 0) { fprintf(stderr, "bench mode needs frames_per_batch * depth to be a multiple of n_distinct\n"); return 2; }
			for (int i = 0; i < 2 * fpb * depth; i++) { pipe.push(raw.data() + fb * (size_t)(i % nd), (size_t)w * 3); while (pipe.pop(out, false)) {} }
			pipe.flush();
			while (pipe.pop(out)) {}
			popped = 64;   // no per-frame lines in bench mode
		}
		const auto t0 = std::chrono::high_resolution_clock::now();
		const long long popped0 = popped;
		for (int i = 0; i < total; i++) {
			if (bench == 2) { pipe.next_frame_buffer(); pipe.commit(); }
			else pipe.push(raw.data() + fb * (size_t)(i % nd), (size_t)w * 3);
			while (pipe.pop(out, false)) { if (popped < 64) report(out); popped++; }
		}
		pipe.flush();
		while (pipe.pop(out)) { if (popped < 64) report(out); popped++; }
		const double sec = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		printf("DONE frames %lld seconds %.6f fps %.1f\n", popped - popped0, sec, (popped - popped0) / sec);
	} catch (const std::exception &e) { fprintf(stderr, "exception: %s\n", e.what()); return 1; }
	return 0;
}
