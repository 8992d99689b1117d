// oracle/cvshim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE, not product code.
//
// A minimal stand-in for <opencv2/opencv.hpp> so that the reference's own hot-path
// functions (ER.cpp / OCR.cpp line ranges, see oracle/build_ref.sh) compile UNMODIFIED
// in a container that has no OpenCV C++ SDK.  Only the subset of cv:: that those
// functions touch is provided (8-bit single-channel Mat with ROI views, Rect algebra,
// Mat /= scalar, bilinear resize).  The arithmetic of the three primitives that decide
// parity is a bit-exact restatement of OpenCV 4.x and is pinned against python cv2 4.13
// in tests/test_cvshim_vs_cv2.py:
//   * Mat /= s     : convertTo with float scale, round-half-even, saturate   
//   * cv::resize   : INTER_LINEAR 8UC1 fixed-point (11-bit coeffs), incl. the exact-2x
//                    INTER_AREA switch                                        
//   * Rect & Rect  : empty intersection -> Rect() (all zero)
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <memory>
#include <vector>
#include <string>
#include <algorithm>
#include <stdexcept>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

namespace cv {

struct Point { int x, y; Point() : x(0), y(0) {} Point(int x_, int y_) : x(x_), y(y_) {} };
struct Size  { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Vec3d { double val[3]; double &operator[](int i) { return val[i]; } };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0]=a; val[1]=b; val[2]=c; val[3]=d; } };

struct Rect {
	int x, y, width, height;
	Rect() : x(0), y(0), width(0), height(0) {}
	Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
	Point br() const { return Point(x + width, y + height); }
	Point tl() const { return Point(x, y); }
	int area() const { return width * height; }
};

inline Rect operator&(const Rect &a, const Rect &b)
{
	int x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
	int x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
	if (x2 <= x1 || y2 <= y1) return Rect();
	return Rect(x1, y1, x2 - x1, y2 - y1);
}

namespace flann { struct Index { Index() {} }; }

class Mat {
public:
	int rows, cols;
	size_t step;
	uchar *data;
	std::shared_ptr<std::vector<uchar> > buf;

	Mat() : rows(0), cols(0), step(0), data(nullptr) {}
	Mat(int r, int c, int /*type*/) { create(r, c); }
	// non-owning view of caller memory (used by the oracle's C wrapper)
	Mat(int r, int c, int /*type*/, void *ext, size_t step_) : rows(r), cols(c), step(step_), data((uchar *)ext) {}

	void create(int r, int c)
	{
		rows = r; cols = c; step = (size_t)c;
		buf = std::make_shared<std::vector<uchar> >((size_t)r * c, (uchar)0);
		data = buf->data();
	}
	static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
	int type() const { return CV_8UC1; }
	size_t total() const { return (size_t)rows * cols; }
	bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

	Mat clone() const
	{
		Mat m(rows, cols, CV_8UC1);
		for (int i = 0; i < rows; i++) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols);
		return m;
	}
	Mat operator()(const Rect &r) const
	{
		Mat m;
		m.rows = r.height; m.cols = r.width; m.step = step; m.buf = buf;
		m.data = data + (size_t)r.y * step + r.x;
		return m;
	}
	uchar *ptr(int i = 0) { return data + (size_t)i * step; }
	const uchar *ptr(int i = 0) const { return data + (size_t)i * step; }
	uchar *ptr(int i, int j) { return data + (size_t)i * step + j; }
	template <typename T> T *ptr(int i = 0) { return (T *)(data + (size_t)i * step); }
	template <typename T> T *ptr(int i, int j) { return (T *)(data + (size_t)i * step) + j; }

	// cv::Mat /= s  ==  convertTo(self, -1, 1/s): float multiply, cvRound (half-even), saturate.
	Mat &operator/=(double s)
	{
		const float a = (float)(1.0 / s);
		for (int i = 0; i < rows; i++) {
			uchar *p = ptr(i);
			for (int j = 0; j < cols; j++) {
				long v = lrintf((float)p[j] * a);
				p[j] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v));
			}
		}
		return *this;
	}
};

// cv::resize(src, dst, dsize) for 8UC1, default INTER_LINEAR.
inline void resize(const Mat &src, Mat &dst, Size dsize)
{
	const int sw = src.cols, sh = src.rows, dw = dsize.width, dh = dsize.height;
	Mat out(dh, dw, CV_8UC1);
	if (dw <= 0 || dh <= 0) { dst = out; return; }
	if (sw == 2 * dw && sh == 2 * dh) {
		// INTER_LINEAR is switched to INTER_AREA for an exact 2x decimation: 2x2 box, round-to-nearest.
		for (int y = 0; y < dh; y++) {
			const uchar *r0 = src.ptr(2 * y), *r1 = src.ptr(2 * y + 1);
			uchar *d = out.ptr(y);
			for (int x = 0; x < dw; x++)
				d[x] = (uchar)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
		}
		dst = out;
		return;
	}
	const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
	std::vector<int> xofs(dw), yofs(dh);
	std::vector<short> ia(2 * dw), ib(2 * dh);
	for (int dx = 0; dx < dw; dx++) {
		float fx = (float)((dx + 0.5) * scale_x - 0.5);
		int sx = (int)floorf(fx);
		fx -= sx;
		if (sx < 0) { fx = 0; sx = 0; }
		if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
		xofs[dx] = sx;
		ia[2 * dx] = (short)lrintf((1.f - fx) * 2048.f);
		ia[2 * dx + 1] = (short)lrintf(fx * 2048.f);
	}
	for (int dy = 0; dy < dh; dy++) {
		float fy = (float)((dy + 0.5) * scale_y - 0.5);
		int sy = (int)floorf(fy);
		fy -= sy;
		yofs[dy] = sy;
		ib[2 * dy] = (short)lrintf((1.f - fy) * 2048.f);
		ib[2 * dy + 1] = (short)lrintf(fy * 2048.f);
	}
	std::vector<int> h0(dw), h1(dw);
	for (int dy = 0; dy < dh; dy++) {
		int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
		int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
		const uchar *r0 = src.ptr(sy0), *r1 = src.ptr(sy1);
		for (int dx = 0; dx < dw; dx++) {
			int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
			h0[dx] = r0[sx] * ia[2 * dx] + r0[sx1] * ia[2 * dx + 1];
			h1[dx] = r1[sx] * ia[2 * dx] + r1[sx1] * ia[2 * dx + 1];
		}
		const int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
		uchar *d = out.ptr(dy);
		for (int dx = 0; dx < dw; dx++) {
			int v = (((b0 * (h0[dx] >> 4)) >> 16) + ((b1 * (h1[dx] >> 4)) >> 16) + 2) >> 2;
			d[dx] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v));
		}
	}
	dst = out;
}

} // namespace cv
