// clust.cu -- MagmaClust entry points (placeholder during bring-up; replaced below).
#include "common.cuh"
#define NOTYET do { mb_set_error(__FILE__, __LINE__, "MagmaClust path not built yet"); return -4; } while (0)
extern "C" int mb_ve_step(mb_context*, int, const mb_kernel*, const double*, const mb_kernel*, const double*, int, const double*, const double*, const double*, int, double, double*, double*, double*) { NOTYET; }
extern "C" int mb_update_mixture(mb_context*, int, const mb_kernel*, const double*, int, const double*, const double*, const double*, double, double*) { NOTYET; }
extern "C" int mb_elbo_clust_tasks(mb_context*, int, const mb_kernel*, const double*, int, const double*, const double*, const double*, double, double*, double*) { NOTYET; }
extern "C" int mb_vm_step(mb_context*, int, const mb_kernel*, double*, const mb_kernel*, double*, int, int, double*, const double*, const double*, const double*, double, double*, int64_t*) { NOTYET; }
extern "C" int mb_elbo_monitoring(mb_context*, int, const mb_kernel*, const double*, const mb_kernel*, const double*, int, const double*, const double*, const double*, const double*, const double*, double, double*) { NOTYET; }
