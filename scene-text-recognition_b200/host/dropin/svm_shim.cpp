// svm_shim.cpp -> libertext_svm.so: libsvm's C ABI (inc/svm.h:76-92) on top of libertext.so.
//
// svm_load_model parses the model on the device side (ert_load_svm), svm_predict_probability densifies the sparse
// svm_node list and calls the batched device path (csrc/svm.cu).  The object handed out IS a `struct svm_model`
// (libsvm's layout, the fields a caller may read are filled: param.svm_type / kernel_type / gamma / probability,
// nr_class, l, label) followed by private state.  The reference calls svm_predict_probability from an OpenMP loop
// on one model (src/ER.cpp:728-735 -> src/OCR.cpp:92): calls on one model are serialised by a mutex.
// Device: environment variable ERTEXT_DEVICE (default 0).  No CPU path: svm_load_model returns NULL (as libsvm does
// on failure, src/svm.cpp:2879) when there is no CUDA device, with the reason in ert_last_error().
#include "../../../include/ertext.h"
#include "../../../include/ertext_svm.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#define SHIM_API extern "C" __attribute__((visibility("default")))

namespace {
struct ShimModel {
	svm_model m;          // must stay first: callers hold `svm_model *`
	uint32_t magic;
	ert_ctx *ctx;
	int dims;
	std::mutex mu;
	std::vector<double> dense;
};
const uint32_t MAGIC = 0x45525453u;   // "ERTS"

ShimModel *shim_of(const svm_model *m)
{
	ShimModel *s = reinterpret_cast<ShimModel *>(const_cast<svm_model *>(m));
	return (s && s->magic == MAGIC) ? s : nullptr;
}
int device_ordinal()
{
	const char *e = getenv("ERTEXT_DEVICE");
	return e ? atoi(e) : 0;
}
} // namespace

extern "C" { __attribute__((visibility("default"))) int libsvm_version = LIBSVM_VERSION; }

SHIM_API svm_model *svm_load_model(const char *model_file_name)
{
	ert_ctx *ctx = ert_create(nullptr, device_ordinal());
	if (!ctx) { fprintf(stderr, "libertext_svm: %s\n", ert_last_error()); return nullptr; }
	if (ert_load_svm(ctx, model_file_name) < 0) { ert_destroy(ctx); return nullptr; }     // libsvm: NULL when the file cannot be read / parsed
	ShimModel *s = new (std::nothrow) ShimModel();
	if (!s) { ert_destroy(ctx); return nullptr; }
	memset(&s->m, 0, sizeof s->m);
	s->magic = MAGIC; s->ctx = ctx; s->dims = ert_svm_dims(ctx);
	s->m.param.svm_type = C_SVC; s->m.param.kernel_type = RBF; s->m.param.gamma = ert_svm_gamma(ctx); s->m.param.probability = 1;
	s->m.nr_class = ert_svm_nr_class(ctx);
	s->m.l = ert_svm_total_sv(ctx);
	s->m.label = (int *)malloc(sizeof(int) * (size_t)s->m.nr_class);
	if (!s->m.label || ert_svm_labels(ctx, s->m.label) < 0) { free(s->m.label); ert_destroy(ctx); delete s; return nullptr; }
	s->m.free_sv = 1;
	return &s->m;
}

SHIM_API int svm_get_svm_type(const svm_model *model) { return model->param.svm_type; }
SHIM_API int svm_get_nr_class(const svm_model *model) { return model->nr_class; }
SHIM_API int svm_get_nr_sv(const svm_model *model) { return model->l; }
SHIM_API void svm_get_labels(const svm_model *model, int *label)
{
	if (model->label) for (int i = 0; i < model->nr_class; i++) label[i] = model->label[i];
}
SHIM_API int svm_check_probability_model(const svm_model *model) { return shim_of(model) ? 1 : 0; }

SHIM_API int svm_predict_probability_batch(const svm_model *model, const svm_node *const *x, int n, double *labels, double *prob_estimates)
{
	ShimModel *s = shim_of(model);
	if (!s || n < 0 || (n && (!x || !labels || !prob_estimates))) return -1;
	if (n == 0) return 0;
	std::lock_guard<std::mutex> lk(s->mu);
	s->dense.assign((size_t)n * s->dims, 0.0);
	for (int i = 0; i < n; i++)
		for (const svm_node *p = x[i]; p->index != -1; ++p) {
			// libsvm adds value^2 of an index no support vector has to EVERY distance; the dense device layout has no
			// column for it.  The reference's features never leave the model's 1800 dimensions (src/OCR.cpp:203-218).
			if (p->index < 0 || p->index >= s->dims) { fprintf(stderr, "libertext_svm: feature index %d outside the model's %d dimensions\n", p->index, s->dims); return -1; }
			s->dense[(size_t)i * s->dims + p->index] = p->value;
		}
	return ert_svm_predict_probability_batch(s->ctx, s->dense.data(), n, labels, prob_estimates);
}

SHIM_API double svm_predict_probability(const svm_model *model, const svm_node *x, double *prob_estimates)
{
	double label = NAN;
	const svm_node *const xs[1] = {x};
	if (svm_predict_probability_batch(model, xs, 1, &label, prob_estimates)) return NAN;
	return label;
}

// svm_predict (src/svm.cpp:2577-2590) votes over the pairwise decision values; with a probability model the reference only
// ever calls svm_predict_probability.  Here: the label of the largest probability estimate.
SHIM_API double svm_predict(const svm_model *model, const svm_node *x)
{
	std::vector<double> p((size_t)(model->nr_class > 0 ? model->nr_class : 1));
	return svm_predict_probability(model, x, p.data());
}

SHIM_API void svm_free_model_content(svm_model *model_ptr)
{
	ShimModel *s = shim_of(model_ptr);
	if (!s) return;
	free(s->m.label); s->m.label = nullptr;
	if (s->ctx) { ert_destroy(s->ctx); s->ctx = nullptr; }
}

SHIM_API void svm_free_and_destroy_model(svm_model **model_ptr_ptr)
{
	if (!model_ptr_ptr || !*model_ptr_ptr) return;
	ShimModel *s = shim_of(*model_ptr_ptr);
	if (s) { svm_free_model_content(&s->m); s->magic = 0; delete s; }
	*model_ptr_ptr = nullptr;
}
