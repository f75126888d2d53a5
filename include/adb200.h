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

#if defined(__GNUC__)
#define ADB_API __attribute__((visibility("default")))
#else
#define ADB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ADB_MAX_RANK 8
#define ADB_MAX_OPERANDS 8
#define ADB_RED_MAX_OPS 4   /* operands of an einsum that still has summed letters and is not a GEMM */

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
ADB_API int adb_init(int device);                    /* cudaSetDevice + driver entry points; call once per process */
ADB_API const char* adb_last_error(void);            /* thread-local message of the last failing call */
ADB_API const char* adb_version(void);
ADB_API int64_t adb_launch_count(void);              /* kernels launched by this library so far (bench.py gpu_launches) */
ADB_API int adb_sync(void* stream);                  /* cudaStreamSynchronize */

/* ---- memory: replaces numpy array ownership (reference node.py:111-129 Variable._value) --------- */
ADB_API int adb_malloc(void** out, size_t bytes, void* stream);          /* stream-ordered pool allocation */
ADB_API int adb_free(void* ptr, void* stream);
ADB_API int adb_memcpy_h2d(void* dst, const void* host_src, size_t bytes, void* stream);
ADB_API int adb_memcpy_d2h(void* host_dst, const void* src, size_t bytes, void* stream);
ADB_API int adb_host_alloc(void** out, size_t bytes);                    /* pinned host staging */
ADB_API int adb_host_free(void* ptr);
/* pitched copies: `width_bytes` per row, `rows` rows; pitches in bytes (host side may be a strided numpy view) */
ADB_API int adb_memcpy2d_h2d(void* dst, size_t dst_pitch, const void* host_src, size_t src_pitch, size_t width_bytes, size_t rows, void* stream);
ADB_API int adb_memcpy2d_d2h(void* host_dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, void* stream);

/* ---- contraction: replaces np.einsum at reference autodiff/core/ops.py:180 ---------------------- */
ADB_API int adb_gemm_f64(const adb_gemm_desc* desc, void* stream);
/* path: 0 = auto, 1 = force TMA+DMMA kernel (error if not describable), 2 = force generic kernel,
 * 3 = force TMA+DMMA with 64x64 tiles, 4 = force TMA+DMMA with 128x128 tiles, 5 = force the 256x32 tall-skinny tiles,
 * 6 = force the 128x32 tiles (few rows and few columns, long K), 7 / 8 = force the stream-K kernel with 128x128 / 64x64 tiles
 * (whole-tile CTAs for the full waves + CTAs sharing the ragged last wave along K, one launch; bitwise repeatable) */
ADB_API int adb_gemm_f64_path(const adb_gemm_desc* desc, void* stream, int path);
/* The stream-K schedule the GEMM would use for tiles of bm x bn on `slots` CTA slots (no GPU needed; CPU tests of the schedule):
 * out[0..5] = k-tiles per tile, tiles, tiles taken whole (full waves), k-tile units of the shared region, CTAs sharing it, grid size. */
ADB_API int adb_gemm_sk_plan(long long M, long long N, long long K, long long batch, int bm, int bn, long long slots, int min_units, long long* out);
/* Same contract (fp64 in, fp64 out), arithmetic as 3xTF32 on the tcgen05 tensor cores with fp32 accumulation in TMEM:
 * a ~= hi + lo, A.B ~= hi.hi + lo.hi + hi.lo  (parity bar 1e-4; measured ~1e-5 at K = 8192: fp32 accumulation chains of <= 2048,
 * chains added in fp64).  Env ADB_TF32_2CTA=0 pins the single-CTA kernel, =2 forces the CTA-pair (cta_group::2) kernel. */
ADB_API int adb_gemm_tf32x3(const adb_gemm_desc* desc, void* stream);
/* Contraction precision used by adb_einsum for GEMM-shaped einsums: 0 = FP64 DMMA (default), 1 = 3xTF32 tcgen05. */
ADB_API int adb_set_precision(int precision);
ADB_API int adb_get_precision(void);


/* ---- fused elementwise programs: replace the ufunc bodies of reference autodiff/core/ops.py
 *      (:53 Add, :75 Mul, :96 Negate, :116 Recipr, :228 ReLU, :274 SigmoidCE, :290 Pow, :309 Log, :339 Exp,
 *      :354 Sigmoid, :369 sqrt, :391 NormalDistribution) and the copies of reshape.py (:35 Concat,
 *      :94 Slice, :99-100 scatter, :126-129 Pad).
 *  Accumulator machine: `acc` plus 8 temps, evaluated per element.  op = KIND | SRC where SRC says where
 *  the second operand x comes from.  Inputs broadcast numpy-style against the (shared) output shape. */
#define ADB_EW_MAX_IN 8
#define ADB_EW_MAX_OUT 4
#define ADB_EW_MAX_INSTR 64
#define ADB_EW_SRC_NONE 0
#define ADB_EW_SRC_IN 1    /* x = input tensor `arg`            */
#define ADB_EW_SRC_IMM 2   /* x = `imm`                          */
#define ADB_EW_SRC_TMP 3   /* x = temp `arg`                     */
enum {
    ADB_EW_LOAD = 0 << 2,       /* acc = x                          */
    ADB_EW_ADD = 1 << 2,        /* acc = acc + x                    */
    ADB_EW_MUL = 2 << 2,        /* acc = acc * x                    */
    ADB_EW_DIV = 3 << 2,        /* acc = acc / x                    */
    ADB_EW_POW = 4 << 2,        /* acc = pow(acc, x)                */
    ADB_EW_RPOW = 5 << 2,       /* acc = pow(x, acc)                */
    ADB_EW_MAX = 6 << 2,        /* acc = fmax(acc, x)               */
    ADB_EW_STORE_T = 7 << 2,    /* temp[arg] = acc                  */
    ADB_EW_STORE_OUT = 8 << 2,  /* out[arg][...] = acc              */
    ADB_EW_NEG = 9 << 2,        /* acc = -acc                       */
    ADB_EW_RECIP_EPS = 10 << 2, /* acc = 1 / (acc + imm)   ops.py:116 */
    ADB_EW_LOG_EPS = 11 << 2,   /* acc = log(acc + imm)    ops.py:309 */
    ADB_EW_EXP = 12 << 2,
    ADB_EW_RELU = 13 << 2,      /* acc = acc * (acc > 0)   ops.py:228 */
    ADB_EW_SIGMOID = 14 << 2,   /* acc = 1 / (1 + exp(-acc)) ops.py:354 */
    ADB_EW_SQRT = 15 << 2,
    ADB_EW_ABS = 16 << 2,
    ADB_EW_SQUARE = 17 << 2,
    ADB_EW_LOG = 18 << 2,
    ADB_EW_RECIP = 19 << 2      /* acc = 1 / acc                    */
};
typedef struct adb_ew_instr {
    int32_t op;
    int32_t arg;
    double imm;
} adb_ew_instr;
ADB_API int adb_ewise(const adb_ew_instr* code, int n_instr, int n_in, const adb_tensor* ins, int n_out,
              const adb_tensor* outs, void* stream);

/* Diagnostics for the program dispatcher: programs listed in csrc/ewise_catalog.txt are compiled to straight-line
 * kernels, everything else runs on the bytecode interpreter.  adb_ewise_stats writes one line per distinct program
 * seen so far ("<calls> <quads> <static|interp> op:arg ...") and returns the buffer size needed;
 * adb_ewise_force_interpreter(1) disables the catalogue (A/B measurements, parity tests of both paths). */
ADB_API int adb_ewise_stats(char* buf, size_t buflen);
ADB_API int adb_ewise_force_interpreter(int on);
/* test hook: quads per launch before an iteration space is split on the host (0 = the real limit, 2^31 - 1) */
ADB_API int adb_ewise_set_piece_limit(long long quads);

/* ---- einsum: replaces np.einsum(self.op_str, *arr) at reference autodiff/core/ops.py:180 ----------
 *  op_str uses single letters [a-zA-Z] with an explicit "->" (the host expands "..." first, mirroring
 *  ops.py:131,167-170).  The planner canonicalises the subscripts into one of: alias-free copy/permute,
 *  broadcast product (no summed letter), reduction (warp-shuffle rows / tiled columns), or
 *  strided-batched FP64 GEMM on DMMA tensor cores; >2 operands with summed letters are contracted by the
 *  reduction kernel directly.  `out` must be preallocated with the einsum's result shape. */
ADB_API int adb_einsum(const char* op_str, int n_ops, const adb_tensor* ops, const adb_tensor* out, void* stream);
/* Host-only: writes a one-line description of the plan adb_einsum would run ("gemm M=.. N=.. K=.. batch=..",
 * "reduce rows ...", "ewise ...") into buf.  Needs no GPU; used by the CPU tests of the planner. */
ADB_API int adb_einsum_describe(const char* op_str, int n_ops, const adb_tensor* ops, const adb_tensor* out, char* buf, size_t buflen);

/* ---- reductions: replace np.sum(x, axis, keepdims=True) at reshape.py:14 (ReduceSumKeepDims) and
 *      np.sum(np.square(x)) at ops.py:369 (FrobeniusNorm); expressed through the same reduction kernels
 *      as single-operand einsums.  out has x's rank with size-1 dims on the reduced axes. */
ADB_API int adb_reduce_sum(const adb_tensor* x, const adb_tensor* out, void* stream);
ADB_API int adb_sum_squares(const adb_tensor* x, double* out_scalar, void* stream);
/* sqrt(np.sum(np.square(x))): FrobeniusNorm of ONE tensor (ops.py:369 with a single node) in one launch - the root is taken by
 * the last slice of the reduction. */
ADB_API int adb_norm2(const adb_tensor* x, double* out_scalar, void* stream);

/* ---- SoftmaxCEWithLogits._eval, reference ops.py:243-255.  labels/logits: (rows, cols) views;
 *      out: rows elements (stride out_stride).  *flag (device int) is OR-ed with 1 when a label row does
 *      not sum to 1 within np.allclose tolerance (ops.py:246-248); the host raises ValueError. */
ADB_API int adb_softmax_ce(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, void* stream);

/* ---- convolution = im2col + einsum.  The reference has only a commented-out design
 *      (reshape.py:137-160 AsStrided windows, high_level_ops.py:152-165 Convolution); semantics from its dead test
 *      tests/test_convolution.py:17-33: NHWC, VALID, stride 1, cross-correlation, filters last.
 *      x: (N,H,W,C) view; cols: (N*OH*OW, kh*kw*C).  col2im is the adjoint (gradient) of im2col. */
ADB_API int adb_im2col(const adb_tensor* x, int kh, int kw, const adb_tensor* cols, void* stream);
ADB_API int adb_col2im(const adb_tensor* cols, int kh, int kw, const adb_tensor* dx, void* stream);

/* ---- implicit-GEMM convolution: the contractions on either side of Im2Col / Col2Im WITHOUT the
 *      materialised im2col matrix (im2col-mode TMA feeds the FP64 DMMA pipeline straight from the NHWC image).  Same values
 *      as the adb_im2col / adb_einsum / adb_col2im sequence they replace (reference design: reshape.py:137-160 windows +
 *      Einsum, high_level_ops.py:152-165).  kh x kw window, stride 1, VALID; C (mode 2: O) must be a multiple of 16.
 *        mode 0  out (N*OH*OW, O)      = Im2Col(x) @ w            x: (N,H,W,C), w: (kh*kw*C, O)
 *        mode 1  out (kh*kw*C, O)      = Im2Col(x)^T @ w          w := the output gradient (N*OH*OW, O)
 *        mode 2  out (N,H,W,C) dense   = Col2Im(x' @ w^T)         x := the output gradient seen as (N,OH,OW,O)
 *      adb_conv_gemm_ok returns 1 when these operands qualify (else the caller runs the three-call sequence). */
ADB_API int adb_conv_gemm(int mode, const adb_tensor* x, int kh, int kw, const adb_tensor* w, const adb_tensor* out, void* stream);
ADB_API int adb_conv_gemm_ok(int mode, const adb_tensor* x, int kh, int kw, const adb_tensor* w, const adb_tensor* out);

/* ---- launch-tape replay as ONE CUDA graph (replaces the per-node Python recursion of reference
 *      node.py:42-46 for graphs that are evaluated again and again).  adb_graph_begin puts `capture_stream` (must not be
 *      the legacy default stream) into relaxed capture mode; every compute call issued on it afterwards is recorded
 *      instead of executed; adb_graph_end instantiates the recording into *exec (0 and an error if a recorded call was
 *      not capturable - the stream always leaves capture mode).  adb_graph_launch replays the whole recording with one
 *      driver call on any stream.  adb_memset_async / adb_memcpy_d2d are the two plain stream operations the replay needs
 *      (re-zeroing the label-check flags inside the graph, staging the leaves next to the graph's arena). */
ADB_API int adb_graph_begin(void* capture_stream);
ADB_API int adb_graph_end(void* capture_stream, void** exec);
ADB_API int adb_graph_launch(void* exec, void* stream);
ADB_API int adb_graph_destroy(void* exec);
ADB_API int adb_memset_async(void* ptr, int byte_value, size_t bytes, void* stream);
ADB_API int adb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADB200_H */
