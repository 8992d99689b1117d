// accumulator_cpu.cpp -- host-only check of FrameAccumulator / VideoTimes (host/FramePipeline.hpp): video_mode's two-frame
// accumulation (src/utils.cpp:96-144: tracked_vec over frame_count frames, channel_vec = the frame with n == frame_count / 2)
// and avg_time[0..3].  No device call is made; prints what the pytest compares.
#include "../../scene-text-recognition_b200/host/FramePipeline.hpp"
#include <cstdio>

static ertx::FrameRegions make(long long idx, int n_cand, const std::vector<int32_t> &tracked)
{
	ertx::FrameRegions f;
	f.frame_index = idx;
	for (int i = 0; i < n_cand; i++) { ert_tracked c; std::memset(&c, 0, sizeof c); c.plane = i % 6; c.x = (int)idx * 100 + i; f.cand.push_back(c); }
	f.tracked = tracked;
	f.extract_ms = 1.0; f.nms_ms = 0.25; f.classify_ms = 0.5; f.track_ms = 0.125;
	return f;
}

int main()
{
	ertx::FrameAccumulator acc(2);
	ertx::VideoTimes times;
	const std::vector<std::vector<int32_t>> tr = {{2, 0}, {1}, {}, {3, 1, 0}, {0}};
	for (int i = 0; i < 5; i++) {
		ertx::FrameRegions f = make(10 + i, 4, tr[(size_t)i]);
		times.add(f);
		const bool ready = acc.push(std::move(f));
		printf("push %d ready %d frames %zu tracked", i, ready ? 1 : 0, acc.frames().size());
		for (const ert_tracked &c : acc.tracked()) printf(" %d", c.x);
		printf(" middle %lld\n", acc.middle_frame_index());
	}
	printf("times %.6f %.6f %.6f %.6f frames %lld\n", times.t[0], times.t[1], times.t[2], times.t[3], times.frames);
	ertx::FrameAccumulator one(1);
	const bool r1 = one.push(make(7, 2, {1}));
	printf("single %d middle %lld\n", r1 ? 1 : 0, one.middle_frame_index());
	return 0;
}
