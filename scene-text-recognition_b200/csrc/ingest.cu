// ingest.cu -- compressed frames in, no host round trip of the pixels.
//
// The reference's video_mode pulls decoded frames out of cv::VideoCapture on the host (src/utils.cpp:107-109 `cap >> frame`)
// and hands them to compute_channels.  Here a batch of JPEG bitstreams is decoded by nvJPEG straight into the context's
// device BGR buffer (interleaved 8UC3, the layout compute_channels reads), on the batch's own stream, and the path continues
// as it does behind ert_detect_classify_device: the decoded pixels never exist in host memory and the 6.2 MB per frame that
// bound the host-input path (the H2D ceiling) shrinks to the ~0.3 MB of the bitstream.
//
// nvJPEG is opened at first use with dlopen (libnvjpeg.so.12 ships with the CUDA toolkit): libertext.so loads without it and
// only this entry point fails, loudly, when it is missing.  Backend: the hardware JPEG engine when the device has one
// (NVJPEG_BACKEND_HARDWARE), else GPU-assisted Huffman (NVJPEG_BACKEND_GPU_HYBRID), else the default backend.
// nvJPEG's IDCT / chroma upsampling are not bit-identical to libjpeg's (what cv::imread / VideoCapture use), so parity for
// this row is stated on the DECODED pixels: ert_jpeg_fetch_frames returns them and the tests feed the same pixels to the oracle.
#include "ctx.h"
#include <dlfcn.h>
#include <nvjpeg.h>

namespace ert {

struct NvjpegApi {
	void *lib = nullptr;
	nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t *, nvjpegPinnedAllocator_t *, unsigned int, nvjpegHandle_t *) = nullptr;
	nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
	nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t *) = nullptr;
	nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
	nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char *, size_t, int *, nvjpegChromaSubsampling_t *, int *, int *) = nullptr;
	nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
	nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char *const *, const size_t *, nvjpegImage_t *, cudaStream_t) = nullptr;
};

static NvjpegApi *nvjpeg_api()
{
	static NvjpegApi api;
	static bool tried = false;
	if (tried) return api.lib ? &api : nullptr;
	tried = true;
	const char *names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
	for (const char *nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (api.lib) break; }
	if (!api.lib) { set_error("nvJPEG is not available: %s", dlerror()); return nullptr; }
#define ERT_NVJ(sym) *(void **)(&api.sym) = dlsym(api.lib, "nvjpeg" #sym); if (!api.sym) { set_error("libnvjpeg lacks nvjpeg" #sym); dlclose(api.lib); api.lib = nullptr; return nullptr; }
	ERT_NVJ(CreateEx) ERT_NVJ(Destroy) ERT_NVJ(JpegStateCreate) ERT_NVJ(JpegStateDestroy) ERT_NVJ(GetImageInfo) ERT_NVJ(DecodeBatchedInitialize) ERT_NVJ(DecodeBatched)
#undef ERT_NVJ
	return &api;
}

struct JpegDecoder {
	nvjpegHandle_t handle = nullptr;
	nvjpegJpegState_t state = nullptr;
	int backend = -1;
	int batch = 0;          // batch size the state is initialised for
	int last_status = 0;
};

static const char *backend_name(int b)
{
	switch (b) {
	case NVJPEG_BACKEND_HARDWARE: return "hardware";
	case NVJPEG_BACKEND_GPU_HYBRID: return "gpu_hybrid";
	case NVJPEG_BACKEND_HYBRID: return "hybrid";
	default: return "default";
	}
}

void jpeg_decoder_destroy(void *p)
{
	JpegDecoder *d = (JpegDecoder *)p;
	if (!d) return;
	NvjpegApi *api = nvjpeg_api();
	if (api) { if (d->state) api->JpegStateDestroy(d->state); if (d->handle) api->Destroy(d->handle); }
	delete d;
}

static JpegDecoder *decoder_of(ert_ctx *c, NvjpegApi *api, int want_backend)
{
	JpegDecoder *d = (JpegDecoder *)c->jpeg;
	if (d && (want_backend < 0 || d->backend == want_backend)) return d;
	if (d) { jpeg_decoder_destroy(d); c->jpeg = nullptr; }
	d = new JpegDecoder;
	const int order[] = {NVJPEG_BACKEND_HARDWARE, NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_DEFAULT};
	for (int b : order) {
		if (want_backend >= 0 && b != want_backend) continue;
		// interpolated chroma upsampling: what libjpeg ("fancy upsampling") and therefore cv::imread / VideoCapture do
		const nvjpegStatus_t rc = api->CreateEx((nvjpegBackend_t)b, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION, &d->handle);
		if (rc == NVJPEG_STATUS_SUCCESS) { d->backend = b; break; }
		d->last_status = (int)rc;
		d->handle = nullptr;
	}
	if (!d->handle) { set_error("nvjpegCreateEx failed (status %d) for %s", d->last_status, want_backend >= 0 ? "the requested backend" : "every backend"); delete d; return nullptr; }
	if (api->JpegStateCreate(d->handle, &d->state) != NVJPEG_STATUS_SUCCESS) { set_error("nvjpegJpegStateCreate failed"); api->Destroy(d->handle); delete d; return nullptr; }
	c->jpeg = d;
	return d;
}

} // namespace ert

using namespace ert;

extern "C" {

int ert_enqueue_jpeg(ert_ctx *c, const uint8_t *const *data, const size_t *sizes, int n_frames, int W, int H, int upto)
{
	if (!c || !data || !sizes || n_frames < 1 || W < 1 || H < 1) { set_error("bad arguments"); return -1; }
	NvjpegApi *api = nvjpeg_api();
	if (!api) return -1;
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	JpegDecoder *d = decoder_of(c, api, c->jpeg_backend);
	if (!d) return -1;
	for (int i = 0; i < n_frames; i++) {
		int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
		nvjpegChromaSubsampling_t sub;
		if (!data[i] || api->GetImageInfo(d->handle, data[i], sizes[i], &ncomp, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS) { set_error("frame %d is not a JPEG bitstream nvJPEG understands", i); return -1; }
		if (ws[0] != W || hs[0] != H) { set_error("frame %d is %dx%d, the batch is declared %dx%d", i, ws[0], hs[0], W, H); return -1; }
	}
	// the previous batch of this context may still be reading d_bgr: the decode is ordered behind it on the same stream
	const size_t frame_bytes = (size_t)W * 3 * H;
	if (c->bgr_cap < frame_bytes * n_frames) {
		ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
		cudaFree(c->d_bgr); c->d_bgr = nullptr; c->bgr_cap = 0;
		ERT_CUDA_CHECK(cudaMalloc((void **)&c->d_bgr, frame_bytes * n_frames));
		c->bgr_cap = frame_bytes * n_frames;
	}
	if (d->batch != n_frames) {
		if (api->DecodeBatchedInitialize(d->handle, d->state, n_frames, c->jpeg_cpu_threads, NVJPEG_OUTPUT_BGRI) != NVJPEG_STATUS_SUCCESS) {
			set_error("nvjpegDecodeBatchedInitialize failed (backend %s, batch %d)", backend_name(d->backend), n_frames);
			return -1;
		}
		d->batch = n_frames;
	}
	std::vector<nvjpegImage_t> dst((size_t)n_frames);
	for (int i = 0; i < n_frames; i++) {
		memset(&dst[i], 0, sizeof(nvjpegImage_t));
		dst[i].channel[0] = c->d_bgr + frame_bytes * i;
		dst[i].pitch[0] = (size_t)W * 3;
	}
	ERT_CUDA_CHECK(cudaEventRecord(c->ev_decode[0], c->stream));
	const nvjpegStatus_t rc = api->DecodeBatched(d->handle, d->state, data, sizes, dst.data(), c->stream);
	if (rc != NVJPEG_STATUS_SUCCESS) { set_error("nvjpegDecodeBatched failed with status %d (backend %s)", (int)rc, backend_name(d->backend)); return -1; }
	ERT_CUDA_CHECK(cudaEventRecord(c->ev_decode[1], c->stream));
	c->jpeg_frames = n_frames; c->jpeg_W = W; c->jpeg_H = H;
	return ert_detect_classify_device(c, c->d_bgr, n_frames, W, H, W * 3, upto);
}

/* the pixels the last ert_enqueue_jpeg of this context decoded ([n_frames][H][W][3] BGR); waits for the context's stream */
int ert_jpeg_fetch_frames(ert_ctx *c, uint8_t *bgr_out)
{
	if (!c || !bgr_out || c->jpeg_frames < 1) { set_error("no decoded batch on this context"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
	ERT_CUDA_CHECK(cudaMemcpy(bgr_out, c->d_bgr, (size_t)c->jpeg_frames * c->jpeg_W * c->jpeg_H * 3, cudaMemcpyDeviceToHost));
	return 0;
}

/* backend: -1 = best available (default), else an nvjpegBackend_t value (3 hardware, 2 gpu_hybrid, 0 default); cpu_threads: nvJPEG's host Huffman pool */
int ert_set_jpeg_backend(ert_ctx *c, int backend, int cpu_threads)
{
	if (!c || backend < -1 || backend > 6 || cpu_threads < 1) { set_error("bad arguments"); return -1; }
	c->jpeg_backend = backend; c->jpeg_cpu_threads = cpu_threads;
	return 0;
}

/* the backend in use after the first ert_enqueue_jpeg ("hardware", "gpu_hybrid", "default"), NULL before */
const char *ert_jpeg_backend_name(ert_ctx *c)
{
	if (!c || !c->jpeg) return nullptr;
	return backend_name(((JpegDecoder *)c->jpeg)->backend);
}

/* device milliseconds of the last batch's decode (events around nvjpegDecodeBatched on the batch's stream); -1 if unavailable */
double ert_jpeg_decode_ms(ert_ctx *c)
{
	if (!c || c->jpeg_frames < 1) return -1.0;
	if (cudaEventSynchronize(c->ev_decode[1]) != cudaSuccess) return -1.0;
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, c->ev_decode[0], c->ev_decode[1]) != cudaSuccess) return -1.0;
	return (double)ms;
}

} // extern "C"
