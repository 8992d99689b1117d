/* ertext_svm.h -- the part of libsvm 3.21's C ABI that libertext_svm.so implements on the B200.
 *
 * The reference binds libsvm through exactly these symbols (inc/svm.h:76-92; call sites src/OCR.cpp:18-21 svm_load_model,
 * src/OCR.cpp:55-56,91-92 svm_get_nr_class + svm_predict_probability).  libertext_svm.so exports them with libsvm's own
 * prototypes and struct layouts (struct svm_node inc/svm.h:11-15, svm_parameter inc/svm.h:26-46, svm_model inc/svm.h:51-74),
 * so a program built against the reference's unmodified inc/svm.h links against it instead of src/svm.cpp and runs
 * svm_predict_probability on the device (kernels of csrc/svm.cu).  This header restates those declarations for
 * callers that do not have the reference's header; tests/test_dropin_cpu.py checks the layouts against inc/svm.h.
 *
 * Not implemented (training side of libsvm, outside the path): svm_train, svm_cross_validation, svm_save_model,
 * svm_predict_values, svm_check_parameter, svm_get_sv_indices, svm_get_svr_probability.
 */
#ifndef _LIBSVM_H
#define _LIBSVM_H
#define LIBSVM_VERSION 321
#ifdef __cplusplus
extern "C" {
#endif

extern int libsvm_version;

struct svm_node { int index; double value; };                      /* terminated by index = -1 */
struct svm_problem { int l; double *y; struct svm_node **x; };
enum { C_SVC, NU_SVC, ONE_CLASS, EPSILON_SVR, NU_SVR };            /* svm_type */
enum { LINEAR, POLY, RBF, SIGMOID, PRECOMPUTED };                  /* kernel_type */

struct svm_parameter {
	int svm_type, kernel_type, degree;
	double gamma, coef0;
	double cache_size, eps, C;
	int nr_weight;
	int *weight_label;
	double *weight;
	double nu, p;
	int shrinking, probability;
};

struct svm_model {
	struct svm_parameter param;
	int nr_class;            /* number of classes */
	int l;                   /* total #SV */
	struct svm_node **SV;    /* NULL here: the support vectors live in device memory */
	double **sv_coef;        /* NULL here */
	double *rho, *probA, *probB;   /* NULL here */
	int *sv_indices;         /* NULL */
	int *label;              /* label of each class (label[k]) */
	int *nSV;                /* NULL here */
	int free_sv;
};

struct svm_model *svm_load_model(const char *model_file_name);
int svm_get_svm_type(const struct svm_model *model);
int svm_get_nr_class(const struct svm_model *model);
void svm_get_labels(const struct svm_model *model, int *label);
int svm_get_nr_sv(const struct svm_model *model);
double svm_predict(const struct svm_model *model, const struct svm_node *x);
double svm_predict_probability(const struct svm_model *model, const struct svm_node *x, double *prob_estimates);
void svm_free_model_content(struct svm_model *model_ptr);
void svm_free_and_destroy_model(struct svm_model **model_ptr_ptr);
int svm_check_probability_model(const struct svm_model *model);

/* extension (not in libsvm): score n sparse vectors in ONE device batch; labels[n], prob_estimates n x nr_class */
int svm_predict_probability_batch(const struct svm_model *model, const struct svm_node *const *x, int n, double *labels, double *prob_estimates);

#ifdef __cplusplus
}
#endif
#endif /* _LIBSVM_H */
