// Kronecker / Tucker path (placeholder translation unit; kernels land in the next commits).
#include <cuda_runtime.h>
#include "../../include/ffgp.h"
namespace ffgp { int fail(int code, const char* fmt, const char* a); }
using ffgp::fail;
extern "C" {
size_t ffgp_kernel_matrix_bwd_scratch_bytes(int, int, int, int) { return 0; }
int ffgp_kernel_matrix_bwd_f64(const double*, const double*, const double*, const double*, const double*, int, int, int, int, int,
                               double*, double*, void*, size_t, void*) { return fail(-99, "not implemented%s", ""); }
int ffgp_mode_dot_f64(const double*, const double*, double*, long long, int, long long, int, int, void*) { return fail(-99, "not implemented%s", ""); }
size_t ffgp_syevj_workspace_bytes(int, int) { return 0; }
int ffgp_syevj_f64(const double*, int, int, double*, double*, void*, size_t, int*, void*) { return fail(-99, "not implemented%s", ""); }
int ffgp_kron_core_f64(const double*, const double*, const int*, int, const double*, double, double*, double*, double*, void*, size_t, void*) { return fail(-99, "not implemented%s", ""); }
size_t ffgp_kron_core_scratch_bytes(long long) { return 0; }
}
