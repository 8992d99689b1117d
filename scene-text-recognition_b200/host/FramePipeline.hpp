// FramePipeline.hpp -- header-only C++ streaming runtime over the C ABI (include/ertext.h).
//
// The reference's video_mode (src/utils.cpp:57-228) grabs a frame, runs compute_channels and the per-channel loop
// er_tree_extract -> non_maximum_supression -> classify (src/utils.cpp:115-138), then er_track (src/utils.cpp:140),
// one frame at a time.  This class keeps that per-frame contract -- frames go in, each frame's strong / weak / tracked
// regions come out in the same order -- but feeds the device the way it wants to be fed: frames are staged in pinned
// memory, grouped into batches, and `depth` contexts (streams) are used round-robin so that the H2D copy and the narrow
// kernels of one batch overlap the tile kernel of another (bench.py's e2e path, in C++).
//
//   ertx::FramePipeline pipe(params, /*device*/0, 1920, 1080, "strong.classifier", "weak.classifier");
//   while (cap.read(frame)) { pipe.push(frame.data, frame.step); while (pipe.pop(out, /*wait=*/false)) consume(out); }
//   pipe.flush(); while (pipe.pop(out)) consume(out);
#pragma once
#include "../../include/ertext.h"

#include <cstring>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

namespace ertx {

// Per-frame result: what video_mode hands to er_grouping / er_ocr.
struct FrameRegions {
	long long frame_index = -1;
	std::vector<ert_node> pool[6];         // non_maximum_supression's pool per channel (pool order)
	std::vector<int32_t> label[6];         // ERT_LABEL_STRONG / WEAK / NONE per pooled region
	std::vector<ert_tracked> cand;         // er_track: every strong then weak region with colour / centre / ch
	int n_strong = 0;
	std::vector<int32_t> tracked;          // `tracked` (all_er) as indices into cand, the reference's order
	double batch_ms = 0;                   // device time of the batch this frame travelled in
	// the batch's device time per module, divided by its frames: the shares video_mode adds into avg_time[0..3]
	// (src/utils.cpp:152-157: extract, nms, classify, track) -- plus decode for compressed input
	double extract_ms = 0, nms_ms = 0, classify_ms = 0, track_ms = 0, decode_ms = 0;
};

// video_mode's avg_time[7] (src/utils.cpp:98, 152-157, 205-211): [0] extract [1] nms [2] classify [3] track [4] grouping
// [5] ocr [6] total wall seconds.  [0..3] are device milliseconds here (the reference's are the slowest channel's host
// seconds), accumulated per frame; [4] and [5] are whatever the caller measures around its er_grouping / OCR step.
struct VideoTimes {
	double t[7] = {0, 0, 0, 0, 0, 0, 0};
	long long frames = 0;
	void add(const FrameRegions &f) { t[0] += f.extract_ms * 1e-3; t[1] += f.nms_ms * 1e-3; t[2] += f.classify_ms * 1e-3; t[3] += f.track_ms * 1e-3; frames++; }
};

// video_mode's outer loop (src/utils.cpp:102-211): the tracked regions of `frame_count` (2) consecutive frames are
// accumulated (tracked_vec), the channels of frame frame_count / 2 are kept for er_ocr (channel_vec), then er_grouping /
// er_ocr run once per group.  push() returns true when a group is complete; the caller hands `tracked` (and the middle
// frame's index) to the reference's own er_grouping, which stays host code outside the path (out of scope).
class FrameAccumulator {
public:
	explicit FrameAccumulator(int frame_count = 2) : frame_count_(frame_count < 1 ? 1 : frame_count) {}
	bool push(FrameRegions &&f)
	{
		if ((int)frames_.size() == frame_count_) clear();
		for (int32_t i : f.tracked) tracked_.push_back(f.cand[(size_t)i]);
		if ((int)frames_.size() == frame_count_ / 2) middle_index_ = f.frame_index;
		frames_.push_back(std::move(f));
		return (int)frames_.size() == frame_count_;
	}
	const std::vector<ert_tracked> &tracked() const { return tracked_; }      // tracked_vec
	long long middle_frame_index() const { return middle_index_; }             // the frame whose channels er_ocr reads
	const std::vector<FrameRegions> &frames() const { return frames_; }
	void clear() { frames_.clear(); tracked_.clear(); middle_index_ = -1; }

private:
	int frame_count_;
	std::vector<FrameRegions> frames_;
	std::vector<ert_tracked> tracked_;
	long long middle_index_ = -1;
};

class FramePipeline {
public:
	FramePipeline(const ert_params &prm, int device, int width, int height, const std::string &strong_path, const std::string &weak_path,
	              int frames_per_batch = 8, int depth = 4, int upto = ERT_STAGE_TRACK)
	    : W_(width), H_(height), fpb_(frames_per_batch), upto_(upto)
	{
		if (width < 1 || height < 1 || frames_per_batch < 1 || depth < 1) throw std::runtime_error("FramePipeline: bad geometry");
		slots_.resize((size_t)depth);
		for (Slot &s : slots_) {
			s.ctx = ert_create(&prm, device);
			if (!s.ctx) fail("ert_create");
			if (ert_load_cascade(s.ctx, ERT_CASCADE_STRONG, strong_path.c_str()) < 0) fail("strong cascade");
			if (ert_load_cascade(s.ctx, ERT_CASCADE_WEAK, weak_path.c_str()) < 0) fail("weak cascade");
			s.staging = (unsigned char *)ert_host_alloc(frame_bytes() * (size_t)fpb_);
			if (!s.staging) fail("ert_host_alloc");
		}
	}
	~FramePipeline()
	{
		for (Slot &s : slots_) { if (s.ctx) ert_destroy(s.ctx); ert_host_free(s.staging); }
	}

	// one 8UC3 BGR frame (rows `stride` bytes apart); copied into the pinned staging buffer of the batch being filled
	void push(const unsigned char *bgr, size_t stride)
	{
		unsigned char *dst = next_frame_buffer();
		if (stride == (size_t)W_ * 3) std::memcpy(dst, bgr, frame_bytes());
		else for (int y = 0; y < H_; y++) std::memcpy(dst + (size_t)y * W_ * 3, bgr + (size_t)y * stride, (size_t)W_ * 3);
		commit();
	}
	// zero-copy form: a decoder writes the next frame (W*H*3 bytes, rows W*3 apart) straight into pinned memory ...
	unsigned char *next_frame_buffer()
	{
		Slot &s = slots_[fill_];
		if (s.in_flight) collect(fill_);                       // all slots busy: take the oldest batch home first
		if (s.n > 0 && s.compressed) throw std::runtime_error("FramePipeline: raw and JPEG frames in one batch");
		return s.staging + frame_bytes() * (size_t)s.n;
	}
	// ... and commits it
	void commit()
	{
		Slot &s = slots_[fill_];
		if (s.n == 0) s.first_index = next_index_;
		s.n++; next_index_++;
		if (s.n == fpb_) submit();
	}
	// compressed input: one baseline-JPEG bitstream of a W x H frame.  The bitstream is kept (it is ~20x
	// smaller than the pixels) and the whole batch is decoded on the device by nvJPEG when it is submitted: no pixel H2D copy.
	// A batch is either all compressed or all raw.
	void push_jpeg(const unsigned char *data, size_t size)
	{
		Slot &s = slots_[fill_];
		if (s.in_flight) collect(fill_);
		if (s.n > 0 && !s.compressed) throw std::runtime_error("FramePipeline: raw and JPEG frames in one batch");
		s.compressed = true;
		s.jpeg.emplace_back(data, data + size);
		commit();
	}
	// submit a partially filled batch (end of stream, or latency matters more than throughput)
	void flush() { if (slots_[fill_].n > 0 && !slots_[fill_].in_flight) submit(); }

	// next frame in push order.  wait = true blocks on the oldest batch in flight; returns false when nothing is left
	// (wait) or nothing is ready without blocking (!wait).
	bool pop(FrameRegions &out, bool wait = true)
	{
		if (ready_.empty()) {
			if (order_.empty()) return false;
			if (!wait && ert_batch_done(slots_[order_.front()].ctx) != 1) return false;
			collect(order_.front());
		}
		out = std::move(ready_.front());
		ready_.pop_front();
		return true;
	}
	size_t frames_in_flight() const
	{
		size_t n = 0;
		for (const Slot &s : slots_) if (s.in_flight) n += (size_t)s.n;
		return n;
	}

private:
	struct Slot {
		ert_ctx *ctx = nullptr;
		unsigned char *staging = nullptr;
		int n = 0;
		bool in_flight = false;
		long long first_index = 0;
		bool compressed = false;
		std::vector<std::vector<unsigned char>> jpeg;   // the batch's bitstreams (push_jpeg)
	};
	int W_, H_, fpb_, upto_;
	std::vector<Slot> slots_;
	size_t fill_ = 0;
	long long next_index_ = 0;
	std::deque<size_t> order_;            // slots in flight, oldest first
	std::deque<FrameRegions> ready_;

	size_t frame_bytes() const { return (size_t)W_ * H_ * 3; }
	[[noreturn]] static void fail(const char *what) { throw std::runtime_error(std::string("FramePipeline: ") + what + ": " + ert_last_error()); }

	void submit()
	{
		Slot &s = slots_[fill_];
		if (s.compressed) {
			std::vector<const uint8_t *> ptr; std::vector<size_t> len;
			for (const auto &j : s.jpeg) { ptr.push_back(j.data()); len.push_back(j.size()); }
			if (ert_enqueue_jpeg(s.ctx, ptr.data(), len.data(), s.n, W_, H_, upto_)) fail("ert_enqueue_jpeg");
		} else if (ert_enqueue_host(s.ctx, s.staging, s.n, W_, H_, W_ * 3, upto_)) fail("ert_enqueue_host");
		s.in_flight = true;
		order_.push_back(fill_);
		fill_ = (fill_ + 1) % slots_.size();
	}

	void collect(size_t k)
	{
		// batches are taken home in submission order (a later batch may finish first; its result waits in its context)
		while (!order_.empty()) {
			const size_t j = order_.front();
			order_.pop_front();
			unpack(slots_[j]);
			if (j == k) break;
		}
	}

	void unpack(Slot &s)
	{
		const ert_result *r = nullptr;
		if (ert_fetch_result(s.ctx, &r)) fail("ert_fetch_result");
		if (r->status) throw std::runtime_error(std::string("FramePipeline: device status: ") + ert_status_string(r->status));
		const ert_track_result *t = nullptr;
		if (upto_ >= ERT_STAGE_TRACK && ert_er_track(s.ctx, &t)) fail("ert_er_track");
		for (int f = 0; f < s.n; f++) {
			FrameRegions fr;
			fr.frame_index = s.first_index + f;
			fr.batch_ms = r->stage_ms[5];
			fr.extract_ms = r->stage_ms[0] / s.n; fr.nms_ms = r->stage_ms[1] / s.n; fr.classify_ms = r->stage_ms[2] / s.n;
			fr.track_ms = t ? t->track_ms / s.n : 0.0;
			fr.decode_ms = s.compressed ? ert_jpeg_decode_ms(s.ctx) / s.n : 0.0;
			for (int ch = 0; ch < 6; ch++) {
				const int p = f * 6 + ch;
				const ert_node *nodes = r->nodes + r->node_offset[p];
				for (int k = r->pool_offset[p]; k < r->pool_offset[p + 1]; k++) {
					fr.pool[ch].push_back(nodes[r->pool_node[k]]);
					fr.label[ch].push_back(r->pool_label[k]);
				}
			}
			if (t) {
				fr.cand.assign(t->cand + t->cand_offset[f], t->cand + t->cand_offset[f + 1]);
				fr.n_strong = t->n_strong[f];
				fr.tracked.assign(t->tracked + t->track_offset[f], t->tracked + t->track_offset[f + 1]);
			}
			ready_.push_back(std::move(fr));
		}
		s.n = 0;
		s.in_flight = false;
		s.compressed = false;
		s.jpeg.clear();
	}
};

} // namespace ertx
