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
// and, for the rows after the detect path (er_track / OCR::chain_run), pinned against cv2 in
// tests/test_oracle_next.py:
//   * threshold(THRESH_OTSU) : getThreshVal_Otsu_8u's double-precision recurrence, then src > t ? maxval : 0
//   * findContours(RETR_LIST, CHAIN_APPROX_NONE) : Suzuki-Abe border following on a zero-padded copy
//   * GaussianBlur(7x7, sigma 0) on 8U : fixed-point kernel {8,28,56,72,56,28,8}/256 per axis, BORDER_REFLECT_101
//   * normalize(NORM_MINMAX, CV_8U)   : convertTo with float scale/shift, one fused multiply-add, round-half-even
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
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795
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

// cv::Rect | cv::Rect: bounding box of both; an empty operand is ignored (OpenCV's operator|=)
inline Rect operator|(const Rect &a, const Rect &b)
{
	if (a.width <= 0 || a.height <= 0) return b;
	if (b.width <= 0 || b.height <= 0) return a;
	const int x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
	const int x2 = std::max(a.x + a.width, b.x + b.width), y2 = std::max(a.y + a.height, b.y + b.height);
	return Rect(x1, y1, x2 - x1, y2 - y1);
}
inline Rect &operator|=(Rect &a, const Rect &b) { a = a | b; return a; }
inline Point operator-(const Point &a, const Point &b) { return Point(a.x - b.x, a.y - b.y); }
inline double norm(const Point &p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

namespace flann { struct Index { Index() {} }; }

class Mat {
public:
	int rows, cols;
	size_t step;
	uchar *data;
	std::shared_ptr<std::vector<uchar> > buf;
	int cn;   // channels: 1 (CV_8UC1) everywhere on the path; 3 only for the BGR / YCrCb frames handed to compute_channels

	Mat() : rows(0), cols(0), step(0), data(nullptr), cn(1) {}
	Mat(int r, int c, int type) { create(r, c, type); }
	// non-owning view of caller memory (used by the oracle's C wrapper)
	Mat(int r, int c, int type, void *ext, size_t step_) : rows(r), cols(c), step(step_), data((uchar *)ext), cn(type == CV_8UC3 ? 3 : 1) {}

	void create(int r, int c, int type = CV_8UC1)
	{
		cn = (type == CV_8UC3) ? 3 : 1;
		rows = r; cols = c; step = (size_t)c * cn;
		buf = std::make_shared<std::vector<uchar> >((size_t)r * c * cn, (uchar)0);
		data = buf->data();
	}
	static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
	int type() const { return cn == 3 ? CV_8UC3 : CV_8UC1; }
	int channels() const { return cn; }
	size_t total() const { return (size_t)rows * cols; }
	bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

	Mat clone() const
	{
		Mat m(rows, cols, type());
		for (int i = 0; i < rows; i++) memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * cn);
		return m;
	}
	Mat operator()(const Rect &r) const
	{
		Mat m;
		m.rows = r.height; m.cols = r.width; m.step = step; m.buf = buf; m.cn = cn;
		m.data = data + (size_t)r.y * step + (size_t)r.x * cn;
		return m;
	}
	uchar *ptr(int i = 0) { return data + (size_t)i * step; }
	const uchar *ptr(int i = 0) const { return data + (size_t)i * step; }
	uchar *ptr(int i, int j) { return data + (size_t)i * step + j; }
	template <typename T> T &at(const Point &p) { return *(T *)(data + (size_t)p.y * step + p.x * sizeof(T)); }
	template <typename T> T &at(int i, int j) { return *(T *)(data + (size_t)i * step + j * sizeof(T)); }
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

// ---- primitives used after the detect path (er_track / OCR::chain_run) ---------------------------------

enum { THRESH_BINARY = 0, THRESH_OTSU = 8 };
enum { RETR_EXTERNAL = 0, RETR_LIST = 1 };
enum { CHAIN_APPROX_NONE = 1 };
enum { NORM_MINMAX = 32 };

// `255 - Mat` (MatExpr): saturating per-element subtraction from a scalar
inline Mat operator-(int s, const Mat &m)
{
	Mat out(m.rows, m.cols, CV_8UC1);
	for (int i = 0; i < m.rows; i++) {
		const uchar *p = m.ptr(i);
		uchar *d = out.ptr(i);
		for (int j = 0; j < m.cols; j++) { int v = s - p[j]; d[j] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
	}
	return out;
}

// OpenCV's getThreshVal_Otsu_8u: the running-mean recurrence in double, strict '>' so the first maximum wins.
inline int otsu_threshold_8u(const Mat &src)
{
	const int N = 256;
	int h[N];
	for (int i = 0; i < N; i++) h[i] = 0;
	for (int i = 0; i < src.rows; i++) { const uchar *p = src.ptr(i); for (int j = 0; j < src.cols; j++) h[p[j]]++; }
	double mu = 0, scale = 1. / (src.cols * src.rows);
	for (int i = 0; i < N; i++) mu += i * (double)h[i];
	mu *= scale;
	double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
	const double feps = 1.1920928955078125e-07;   // FLT_EPSILON
	for (int i = 0; i < N; i++) {
		double p_i = h[i] * scale, q2, mu2, sigma;
		mu1 *= q1;
		q1 += p_i;
		q2 = 1. - q1;
		if (std::min(q1, q2) < feps || std::max(q1, q2) > 1. - feps) continue;
		mu1 = (mu1 + i * p_i) / q1;
		mu2 = (mu - q1 * mu1) / q2;
		sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
		if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
	}
	return (int)max_val;
}

inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type)
{
	CV_Assert((type & 7) == THRESH_BINARY);
	int t = (type & THRESH_OTSU) ? otsu_threshold_8u(src) : (int)std::floor(thresh);
	int mv = (int)lrint(maxval); mv = mv < 0 ? 0 : (mv > 255 ? 255 : mv);
	Mat out(src.rows, src.cols, CV_8UC1);
	for (int i = 0; i < src.rows; i++) {
		const uchar *p = src.ptr(i);
		uchar *d = out.ptr(i);
		for (int j = 0; j < src.cols; j++) d[j] = p[j] > t ? (uchar)mv : 0;
	}
	dst = out;
	return (double)t;
}

// findContours(image, contours, RETR_LIST, CHAIN_APPROX_NONE): Suzuki-Abe border following as OpenCV runs it.
// Non-zero = foreground; works on a zero-padded signed copy: 1 = untouched foreground, 2 = visited border pixel,
// 2|-128 = visited with the border passing on its right (east) side.  Direction codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE
// (N = smaller y); the trace looks for the next pixel counter-clockwise starting after the one it came from.
inline void findContours(const Mat &image, std::vector<std::vector<Point> > &contours, int mode, int method)
{
	CV_Assert(mode == RETR_LIST && method == CHAIN_APPROX_NONE);
	contours.clear();
	const int W = image.cols + 2, H = image.rows + 2;
	std::vector<signed char> buf((size_t)W * H, 0);
	for (int y = 0; y < image.rows; y++) {
		const uchar *p = image.ptr(y);
		for (int x = 0; x < image.cols; x++) buf[(size_t)(y + 1) * W + x + 1] = p[x] ? 1 : 0;
	}
	const int dx[8] = { 1, 1, 0, -1, -1, -1, 0, 1 }, dy[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
	int delta[16];
	for (int k = 0; k < 16; k++) delta[k] = dy[k & 7] * W + dx[k & 7];
	const int nbd = 2;
	for (int y = 1; y < H - 1; y++) {
		signed char *row = &buf[(size_t)y * W];
		int prev = 0;
		for (int x = 1; x < W; x++) {
			int p = row[x];
			if (p == prev) continue;
			bool is_hole = false;
			bool start = false;
			if (prev == 0 && p == 1) start = true;                 // outer border
			else if (p == 0 && prev >= 1) { start = true; is_hole = true; }   // hole border, seen from its left pixel
			if (start) {
				signed char *i0 = row + x - (is_hole ? 1 : 0);
				Point pt(x - (is_hole ? 1 : 0) - 1, y - 1);
				std::vector<Point> c;
				int s_end = is_hole ? 0 : 4, s = s_end;
				signed char *i1;
				do { s = (s - 1) & 7; i1 = i0 + delta[s]; } while (*i1 == 0 && s != s_end);
				if (s == s_end) {
					*i0 = (signed char)(nbd | -128);
					c.push_back(pt);
				} else {
					signed char *i3 = i0, *i4 = nullptr;
					for (;;) {
						s_end = s;
						while (s < 15) { i4 = i3 + delta[++s]; if (*i4 != 0) break; }
						s &= 7;
						if ((unsigned)(s - 1) < (unsigned)s_end) *i3 = (signed char)(nbd | -128);
						else if (*i3 == 1) *i3 = (signed char)nbd;
						c.push_back(pt);
						pt.x += dx[s]; pt.y += dy[s];
						if (i4 == i0 && i3 == i1) break;
						i3 = i4;
						s = (s + 4) & 7;
					}
				}
				contours.push_back(c);
				p = row[x];
			}
			prev = p;
		}
	}
}

// GaussianBlur(src, dst, Size(7,7), 0) for 8U: OpenCV's fixed-point path; sigma 0 and ksize 7 select the
// tabulated kernel {1/32, 7/64, 7/32, 9/32, ...} = {8,28,56,72,56,28,8}/256; BORDER_REFLECT_101.
inline void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigma)
{
	CV_Assert(ksize.width == 7 && ksize.height == 7 && sigma == 0);
	static const int K[7] = { 8, 28, 56, 72, 56, 28, 8 };
	const int w = src.cols, h = src.rows;
	CV_Assert(w >= 4 && h >= 4);
	std::vector<int> hp((size_t)w * h);
	for (int y = 0; y < h; y++) {
		const uchar *p = src.ptr(y);
		for (int x = 0; x < w; x++) {
			int s = 0;
			for (int k = 0; k < 7; k++) { int xi = x + k - 3; if (xi < 0) xi = -xi; if (xi >= w) xi = 2 * w - 2 - xi; s += K[k] * p[xi]; }
			hp[(size_t)y * w + x] = s;
		}
	}
	Mat out(h, w, CV_8UC1);
	for (int y = 0; y < h; y++) {
		uchar *d = out.ptr(y);
		for (int x = 0; x < w; x++) {
			int s = 0;
			for (int k = 0; k < 7; k++) { int yi = y + k - 3; if (yi < 0) yi = -yi; if (yi >= h) yi = 2 * h - 2 - yi; s += K[k] * hp[(size_t)yi * w + x]; }
			d[x] = (uchar)((s + 32768) >> 16);
		}
	}
	dst = out;
}

// normalize(src, dst, 0, 255, NORM_MINMAX, CV_8U): min/max in double, then convertTo(8U) with float alpha/beta
// evaluated as ONE fused multiply-add per pixel (what cv2 4.13's AVX2 build does; pinned in the tests).
inline void normalize(const Mat &src, Mat &dst, double a, double b, int norm_type, int /*dtype*/)
{
	CV_Assert(norm_type == NORM_MINMAX);
	double smin = 255, smax = 0;
	for (int i = 0; i < src.rows; i++) { const uchar *p = src.ptr(i); for (int j = 0; j < src.cols; j++) { smin = std::min(smin, (double)p[j]); smax = std::max(smax, (double)p[j]); } }
	const double dmin = std::min(a, b), dmax = std::max(a, b);
	const double scale = (dmax - dmin) * (smax - smin > 2.220446049250313e-16 ? 1. / (smax - smin) : 0);
	const double shift = dmin - smin * scale;
	const float fa = (float)scale, fb = (float)shift;
	Mat out(src.rows, src.cols, CV_8UC1);
	for (int i = 0; i < src.rows; i++) {
		const uchar *p = src.ptr(i);
		uchar *d = out.ptr(i);
		for (int j = 0; j < src.cols; j++) { long v = lrintf(fmaf((float)p[j], fa, fb)); d[j] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
	}
	dst = out;
}

} // namespace cv
