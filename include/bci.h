/*
 * bci.h — C ABI of libbci.so, the B200 (sm_100a) executor backend for the
 * `/v1/execute` hot path of bee-code-interpreter.
 *
 * The reference has no FFI for this path (no native numerics at all, see
 * SURVEY.md §2): the arithmetic lives in numpy, reached from agent code that the
 * reference runs as a subprocess (executor/server.rs:149-169), and file ids are
 * made in services/storage.py:44-63.  Each entry point below names the reference
 * interface (file:line) whose work it takes over.  Everything is plain C:
 * pointers + sizes, no exceptions, no torch types.
 *
 * Conventions
 *  - One process drives ONE GPU (`bci_init(device)`); multi-GPU = one process per
 *    GPU, requests sharded round-robin by the host.
 *  - Every op is batch-first: `B` independent requests are packed into ONE launch
 *    per op class.  `in[i]`, `out[i]`, `n[i]` describe request i.  The pointer
 *    ARRAYS live in host memory; the pointers in them are DEVICE pointers
 *    (from bci_alloc, or any cudaMalloc'ed memory of the same device).
 *  - Return value: BCI_OK (0) or a BCI_ERR_* code; `bci_last_error()` gives the
 *    thread-local message.  Nothing here ever falls back to the CPU: without a
 *    usable GPU every entry returns BCI_ERR_CUDA / BCI_ERR_NOT_INIT.
 *  - Ops are enqueued on the library's stream.  Entries that return results into
 *    HOST memory (`bci_reduce`, `bci_argext`, `bci_sha256`, `bci_d2h`) synchronise
 *    before returning; all others are asynchronous until `bci_sync()`.
 *  - All entries are thread-safe (one submission mutex) and never touch Python:
 *    call them with the GIL released.
 */
#ifndef BCI_H
#define BCI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCI_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------- */
#define BCI_OK 0
#define BCI_ERR_INVALID 1     /* bad argument (NULL, B<=0, unknown op/dtype) */
#define BCI_ERR_UNSUPPORTED 2 /* valid numpy call this backend does not take: caller falls back to numpy */
#define BCI_ERR_CUDA 3        /* CUDA runtime/driver error, message in bci_last_error() */
#define BCI_ERR_OOM 4         /* device allocation failed */
#define BCI_ERR_NOT_INIT 5    /* bci_init() has not succeeded in this process */

/* ---- element types (numpy dtype it mirrors) ------------------------------ */
#define BCI_F32 0 /* numpy.float32 */
#define BCI_F64 1 /* numpy.float64 */
#define BCI_I32 2 /* numpy.int32   */
#define BCI_I64 3 /* numpy.int64   */

/* ---- elementwise ops: numpy ufunc each one mirrors ----------------------- */
/* binary (in1[i] != NULL: array∘array; in1 == NULL or in1[i] == NULL: array∘scalar[i]) */
#define BCI_EW_ADD 0      /* numpy.add       */
#define BCI_EW_SUB 1      /* numpy.subtract  */
#define BCI_EW_MUL 2      /* numpy.multiply  */
#define BCI_EW_DIV 3      /* numpy.divide (true_divide; float dtypes only) */
#define BCI_EW_MAXIMUM 4  /* numpy.maximum (NaN-propagating) */
#define BCI_EW_MINIMUM 5  /* numpy.minimum (NaN-propagating) */
#define BCI_EW_RSUB 6     /* scalar - array  */
#define BCI_EW_RDIV 7     /* scalar / array  (float dtypes only) */
/* unary */
#define BCI_EW_SQUARE 32  /* numpy.square    */
#define BCI_EW_ABS 33     /* numpy.absolute  */
#define BCI_EW_NEG 34     /* numpy.negative  */
#define BCI_EW_SQRT 35    /* numpy.sqrt  (float dtypes only) */
#define BCI_EW_EXP 36     /* numpy.exp   (float dtypes only) */
#define BCI_EW_LOG 37     /* numpy.log   (float dtypes only) */
#define BCI_EW_SIN 38     /* numpy.sin   (float dtypes only) */
#define BCI_EW_COS 39     /* numpy.cos   (float dtypes only) */
#define BCI_EW_TANH 40    /* numpy.tanh  (float dtypes only) */
#define BCI_EW_FLOOR 41   /* numpy.floor (float dtypes only; exact) */
#define BCI_EW_CEIL 42    /* numpy.ceil  (float dtypes only; exact) */
#define BCI_EW_LOG1P 43   /* numpy.log1p (float dtypes only) */
#define BCI_EW_EXPM1 44   /* numpy.expm1 (float dtypes only) */

/* ---- full reductions (axis=None) ----------------------------------------- */
#define BCI_RED_SUM 0   /* numpy.sum   : f32->f32, f64->f64, i32->i64, i64->i64 */
#define BCI_RED_MEAN 1  /* numpy.mean  : f32->f32, f64->f64, i32/i64->f64        */
#define BCI_RED_MAX 2   /* numpy.max   : same dtype, NaN-propagating             */
#define BCI_RED_MIN 3   /* numpy.min   : same dtype, NaN-propagating             */
/* pre-transforms fused into the reduction's load (the single-use temporary of
 * `numpy.sum(numpy.square(x))`, examples/benchmark-numpy.py:21, never reaches HBM) */
#define BCI_PRE_NONE 0
#define BCI_PRE_SQUARE 1
#define BCI_PRE_ABS 2

#define BCI_ARGMAX 0 /* numpy.argmax: first occurrence, first NaN wins */
#define BCI_ARGMIN 1 /* numpy.argmin */

typedef union bci_scalar {
  double f;  /* float dtypes: value is cast to the array dtype (numpy weak-scalar rule) */
  int64_t i; /* integer dtypes */
} bci_scalar;

typedef struct bci_devinfo {
  int device;          /* CUDA ordinal this process is bound to */
  int sm_count;        /* 148 on B200 */
  int cc_major, cc_minor;
  size_t total_mem;    /* bytes of HBM */
  char name[64];
} bci_devinfo;

/* ---- runtime -------------------------------------------------------------- */
/* Binds the process to CUDA device `device` (>=0), creates the op stream, the
 * stream-ordered memory pool and the staging rings.  Idempotent for the same
 * device.  Replaces: pod acquisition in kubernetes_code_executor.py:248-264
 * (the "fresh pod from a warm queue" becomes a warm device context). */
int bci_init(int device);
void bci_shutdown(void);
int bci_abi_version(void);
int bci_device_count(int* count);
int bci_device_info(bci_devinfo* info);
int bci_mem_info(size_t* free_bytes, size_t* total_bytes);
const char* bci_last_error(void);
/* 0 once a CUDA error has left the context unusable (illegal address, launch failure, ECC...): the
 * owner process should exit and be restarted — the executor does that for its broker, the way the
 * reference discards a pod that failed (kubernetes_code_executor.py:176-182,262-264). */
int bci_healthy(void);

/* Device buffers (256-byte aligned, stream-ordered pool; freed memory is reused
 * by the next wave of requests without returning to the driver). */
int bci_alloc(size_t bytes, void** dptr);
int bci_free(void* dptr);
/* Pinned host staging memory for callers that copy request payloads in/out. */
int bci_host_alloc(size_t bytes, void** hptr);
int bci_host_free(void* hptr);
/* Copies on the op stream.  h2d/d2d are asynchronous when `src` is pinned;
 * d2h synchronises before returning. */
int bci_h2d(void* dst_dev, const void* src_host, size_t bytes);
int bci_d2h(void* dst_host, const void* src_dev, size_t bytes);
int bci_d2d(void* dst_dev, const void* src_dev, size_t bytes);
int bci_memset(void* dst_dev, int byte, size_t bytes);
/* Pinned transfer path between an open file and device memory (the request's `files` map going
 * in, kubernetes_code_executor.py:100-113; output files coming out, :126-142): the bytes pass
 * through page-locked bounce buffers on a dedicated copy stream — pread/pwrite of chunk k+1
 * overlaps the DMA of chunk k — and the library lock is NOT held while data moves, so uploads
 * do not stall other requests' launches.  The descriptor is only pread/pwritten (never mapped:
 * a file truncated underneath cannot raise SIGBUS); short files fail with BCI_ERR_INVALID.
 * Both calls are ordered after everything already enqueued and return when the copy is done. */
int bci_h2d_fd(void* dst_dev, int fd, uint64_t file_offset, size_t bytes);
int bci_d2h_fd(int fd, uint64_t file_offset, const void* src_dev, size_t bytes);
int bci_sync(void);

/* Timing on the op stream with CUDA events (what bench.py reports). */
int bci_timer_begin(void);
int bci_timer_end(float* elapsed_ms); /* synchronises */
/* Event slots (0..63) on the op stream, for timing individual launches inside a step. */
int bci_event_record(int slot);
int bci_event_elapsed(int slot_begin, int slot_end, float* elapsed_ms); /* synchronises on slot_end */
/* Number of kernels this library has launched since bci_init (bench `gpu_launches`). */
uint64_t bci_launch_count(void);
/* Writes > L2-size bytes of scratch so the next timed launch starts cold. */
int bci_flush_l2(void);

/* ---- K1 elementwise (numpy ufuncs reached from agent code) */
int bci_ewise(int op, int dtype, int B, const void* const* in0, const void* const* in1,
              const bci_scalar* scalar, void* const* out, const size_t* n);

/* ---- K2 full reductions; host_out = B slots of 8 bytes (result in the low bytes) */
int bci_reduce(int op, int pre, int dtype, int B, const void* const* in, const size_t* n,
               void* host_out);
/* Fused binary-producer reductions used by the C3 request mix:
 *   host_out[i] = reduce_op( pre( ew_op(in0[i], in1[i] or scalar[i]) (+ addend[i]) ) )
 * e.g. mean(x*y+1.0): ew=MUL, addend=1.0, op=MEAN; max(abs(x-y)): ew=SUB, pre=ABS, op=MAX.
 * Each intermediate is rounded exactly as numpy's separate ufunc calls round it. */
int bci_reduce_fused(int op, int pre, int ew_op, int dtype, int B, const void* const* in0,
                     const void* const* in1, const bci_scalar* scalar, const bci_scalar* addend,
                     const size_t* n, void* host_out);

/* ---- K6 argmax/argmin (first occurrence, NaN first); host_out[i] is int64 --- */
int bci_argext(int op, int dtype, int B, const void* const* in, const size_t* n,
               int64_t* host_out);

/* ---- K7 cumulative sum (numpy.cumsum, 1-D): i32 widens to i64 like numpy ---- */
int bci_cumsum(int dtype, int B, const void* const* in, void* const* out, const size_t* n);

/* ---- K5 sort (numpy.sort) and stable argsort (numpy.argsort(kind="stable")) --
 * bci_sort accepts out[i] == in[i] (ndarray.sort, in place); other overlaps are not allowed. */
int bci_sort(int dtype, int B, const void* const* in, void* const* out, const size_t* n);
int bci_argsort(int dtype, int B, const void* const* in, int64_t* const* out_idx,
                const size_t* n);

/* ---- K8 uniform f64 in [0,1): Philox4x32-10, key = seed[i], counter = element/2.
 * Replaces the *unseeded* numpy.random.rand of examples/benchmark-numpy.py:20
 * (same 53-bit construction and distribution, not the MT19937 stream). */
int bci_rand_f64(int B, void* const* out, const size_t* n, const uint64_t* seed);

/* ---- K8c numpy.random.normal / randn / standard_normal, unseeded: Box-Muller on the Philox stream
 * (third counter word 1), out = loc[i] + scale[i] * z.  Same distribution as numpy's legacy
 * normal() (examples/using_imports.py:21-22 draws normals), not its MT19937 polar-method stream. */
int bci_rand_normal_f64(int B, void* const* out, const size_t* n, const uint64_t* seed,
                        const double* loc, const double* scale);

/* ---- K8b the SEEDED legacy generator, bit-exact: numpy.random.seed(s); numpy.random.rand(n)
 * (examples/using_imports.py:20-22 seeds the global MT19937).  states = B x 625 uint32 in host
 * memory, each numpy's get_state(): key[624] followed by pos (0..624); on return it is the
 * state after drawing n[i] doubles, ready for numpy's set_state().  out[i]: n[i] float64. */
#define BCI_MT19937_STATE_WORDS 625
int bci_rand_mt19937_f64(int B, void* const* out, const size_t* n, uint32_t* states);

/* ---- K8d numpy.random.Generator(PCG64) streams, bit-exact with numpy: the generators BASELINE
 * configs C2/C3/C5 draw their inputs from (`numpy.random.default_rng(seed)`) and
 * C4's payload (`.bytes(n)`).  gen[i] is the bit generator's state as numpy reports it
 * (`bit_generator.state["state"]`: 128-bit `state` and `inc`); out[i] receives n[i] elements that
 * numpy would have produced starting from that state.  The caller advances the host generator
 * (PCG64.advance) by the draws consumed: n for F64, ceil(n/2) for the 32-bit kinds.
 *   BCI_PCG_F64            Generator.random(n)                  (next64 >> 11) * 2^-53
 *   BCI_PCG_F32            Generator.random(n, dtype=float32)   (next32 >> 8)  * 2^-24
 *   BCI_PCG_U32            the next32 stream itself             (Generator.bytes, full-range uint32)
 *   BCI_PCG_BOUNDED32_I64  Generator.integers(low, low+range, n)          2 <= range < 2^32, int64 out
 *   BCI_PCG_BOUNDED32_I32  Generator.integers(..., dtype=int32)           int32 out
 * Bounded kinds use numpy's Lemire multiply-shift; a draw numpy would REJECT shifts the rest of
 * its stream, which a parallel generator cannot follow: rejected[i] (host, B words, required for
 * bounded kinds) counts them and the caller falls back to numpy when it is not 0 (it is always
 * 0 for power-of-two ranges).  Asynchronous unless `rejected` is filled. */
#define BCI_PCG_F64 0
#define BCI_PCG_F32 1
#define BCI_PCG_U32 2
#define BCI_PCG_BOUNDED32_I64 3
#define BCI_PCG_BOUNDED32_I32 4
typedef struct bci_pcg64_gen {
  uint64_t state_lo, state_hi; /* bit_generator.state["state"]["state"] */
  uint64_t inc_lo, inc_hi;     /* bit_generator.state["state"]["inc"]   */
  int64_t low;                 /* bounded kinds: lowest value returned  */
  uint64_t range;              /* bounded kinds: high - low             */
} bci_pcg64_gen;
int bci_rand_pcg64(int kind, int B, void* const* out, const size_t* n, const bci_pcg64_gen* gen,
                   uint32_t* rejected);

/* ---- K1+K2+K8 fused: reduce_op(pre(rand(n))) with the uniform stream of bci_rand_f64(seed)
 * generated in registers — `numpy.sum(numpy.square(numpy.random.rand(n)))`, the whole body of
 * examples/benchmark-numpy.py:20-21, without the n*8-byte array ever reaching HBM.  Same values
 * and same summation order as bci_rand_f64 followed by bci_reduce (bit-identical result).
 * host_out = B slots of 8 bytes (float64). */
int bci_reduce_rand_f64(int op, int pre, int B, const size_t* n, const uint64_t* seed, void* host_out);

/* ---- array creation (numpy.zeros/ones/full, numpy.arange, numpy.linspace), SURVEY.md §8a K1:
 * arrays born on the device never cross PCIe.
 *   BCI_FILL_CONST     out[j] = a
 *   BCI_FILL_ARANGE    out[j] = a + j*b   rounded like numpy's @TYPE@_fill (product, then sum);
 *                      a = first element, b = second - first, both already in the array dtype;
 *                      out[1] = c if `c` is given (numpy stores start+step itself)
 *   BCI_FILL_LINSPACE  float64 only: out[j] = j*b + a, out[n-1] = c if `c` is given
 *                      (numpy/_core/function_base.py linspace: arange(num)*step + start, endpoint fix-up) */
#define BCI_FILL_CONST 0
#define BCI_FILL_ARANGE 1
#define BCI_FILL_LINSPACE 2
int bci_fill(int kind, int dtype, int B, void* const* out, const size_t* n, const bci_scalar* a,
             const bci_scalar* b, const bci_scalar* c);

/* ---- numpy.ndarray.astype between the device dtypes (value-preserving or round-to-nearest
 * casts: f32<->f64, i32<->i64, int->float; float->int is left to numpy). */
int bci_cast(int src_dtype, int dst_dtype, int B, const void* const* in, void* const* out, const size_t* n);

/* ---- K3 numpy.matmul for 2-D float32, C[i] = A[i] (MxK) @ Bm[i] (KxN), row-major.
 * 3xTF32 split on tcgen05 with fp32 accumulation in TMEM (error ~1e-7 normwise). */
int bci_matmul_f32(int B, const void* const* A, const void* const* Bm, void* const* C,
                   int M, int N, int K);

/* ---- K4 SHA-256 (FIPS 180-4) of B messages in one launch.
 * Makes services/storage.py:35-39's "identified by their SHA-256 hash" true
 * (the reference uses secrets.token_hex(32), storage.py:52).
 * data[i]: DEVICE pointers if on_device != 0, else HOST pointers (copied in).
 * digests: HOST buffer of B*32 bytes. */
int bci_sha256(int B, const void* const* data, const size_t* len, int on_device,
               uint8_t* digests);
/* Streaming form for the chunked ObjectWriter (storage.py:28-31,
 * kubernetes_code_executor.py:126-134): B writers each absorb one chunk per call. */
typedef struct bci_sha256_state {
  uint32_t h[8];
  uint64_t total_len; /* bytes absorbed so far */
  uint32_t buf_len;   /* bytes waiting in buf (< 64) */
  uint8_t buf[64];
} bci_sha256_state;
int bci_sha256_init(bci_sha256_state* st);
int bci_sha256_update_batch(int B, bci_sha256_state* const* st, const void* const* host_data,
                            const size_t* len);
int bci_sha256_final_batch(int B, bci_sha256_state* const* st, uint8_t* digests);

#ifdef __cplusplus
}
#endif
#endif /* BCI_H */
