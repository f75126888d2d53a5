/* adb200.h — C ABI of libadb200.so, the B200-native execution backend for bgavran/autodiff's
 * node-evaluation hot path.
 *
 * The reference has no FFI: its "plugin interface" is the set of numpy calls inside the `_eval`
 * bodies.  Every entry point below replaces one of those call sites and says which
 * (file:line relative to the reference root).  All pointers are DEVICE pointers unless named host_*;
 * all strides are in ELEMENTS (fp64); every call is asynchronous on `stream` (a cudaStream_t passed as
 * void*) and returns 0 on success, non-zero on error with the message available from
 * adb_last_error().  No torch types cross this boundary.
 */
#ifndef ADB200_H
#define ADB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADB_MAX_RANK 8
#define ADB_MAX_OPERANDS 8

/* A strided fp64 tensor view.  rank 0 = scalar (one element at ptr). Stride 0 = broadcast. */
typedef struct adb_tensor {
    double* ptr;
    int32_t rank;
    int32_t _pad;
    int64_t shape[ADB_MAX_RANK];
    int64_t stride[ADB_MAX_RANK];
} adb_tensor;

/* C[b][m][n] = sum_k A[b][m][k] * B[b][k][n], arbitrary element strides. */
typedef struct adb_gemm_desc {
    int64_t M, N, K, batch;
    const double* A; int64_t a_sm, a_sk, a_sb;
    const double* B; int64_t b_sk, b_sn, b_sb;
    double* C;       int64_t c_sm, c_sn, c_sb;
} adb_gemm_desc;

/* ---- library ---------------------------------------------------------------------------------- */
int adb_init(int device);                    /* cudaSetDevice + driver entry points; call once per process */
const char* adb_last_error(void);            /* thread-local message of the last failing call */
const char* adb_version(void);
int64_t adb_launch_count(void);              /* kernels launched by this library so far (bench.py gpu_launches) */
int adb_sync(void* stream);                  /* cudaStreamSynchronize */

/* ---- memory: replaces numpy array ownership (reference node.py:111-129 Variable._value) --------- */
int adb_malloc(void** out, size_t bytes, void* stream);          /* stream-ordered pool allocation */
int adb_free(void* ptr, void* stream);
int adb_memcpy_h2d(void* dst, const void* host_src, size_t bytes, void* stream);
int adb_memcpy_d2h(void* host_dst, const void* src, size_t bytes, void* stream);
int adb_host_alloc(void** out, size_t bytes);                    /* pinned host staging */
int adb_host_free(void* ptr);

/* ---- contraction: replaces np.einsum at reference autodiff/core/ops.py:180 ---------------------- */
int adb_gemm_f64(const adb_gemm_desc* desc, void* stream);
/* path: 0 = auto, 1 = force TMA+DMMA kernel (error if not describable), 2 = force generic kernel */
int adb_gemm_f64_path(const adb_gemm_desc* desc, void* stream, int path);

#ifdef __cplusplus
}
#endif
#endif /* ADB200_H */
