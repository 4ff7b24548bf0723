/*
 * bci_broker.h — the per-GPU broker of libbci.so and its wire protocol.
 *
 * Each reference request runs as its own OS process (executor/server.rs:152-157; one pod per
 * request, kubernetes_code_executor.py:248-264).  Those sandboxed script processes hold no
 * CUDA context: their numpy proxies send fixed-size op descriptors over a Unix stream socket
 * to the one process per GPU that called bci_init().  The broker
 *   - scopes every device buffer to the connection that created it and frees what a script
 *     leaves behind when it disconnects (the analogue of deleting the pod, :262-264);
 *   - queues compute ops from all connections and drains the queue in a dispatcher thread:
 *     pending ops of the same class become ONE batched call of the C ABI in bci.h, i.e. one
 *     launch per op class for all concurrent requests.
 * Native (C++ threads inside libbci.so): no Python on the per-op path.
 */
#ifndef BCI_BROKER_H
#define BCI_BROKER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- service control (call bci_init() first) -------------------------------------------- */
/* Starts listening on `socket_path` (AF_UNIX, SOCK_STREAM) and returns; ops pending for
 * `coalesce_us` microseconds after the first arrival are packed into the same launches. */
int bci_broker_start(const char* socket_path, int coalesce_us);
void bci_broker_stop(void);
/* stats[0..3] = jobs, batches (launch groups), largest batch, connections accepted */
int bci_broker_stats(uint64_t stats[4]);

/* ---- wire protocol ------------------------------------------------------------------------
 * request  = bci_wire_req  (96 bytes, little endian) + payload_len bytes
 * reply    = bci_wire_rep  (16 bytes)               + payload_len bytes
 * BRK_FREE has no reply.  status is a BCI_* code of bci.h; on error the payload is the
 * message text.  Device buffers are named by their device pointer; a connection may only
 * name buffers it created, within their size.  Sizes are checked for overflow (n * itemsize,
 * M*K*4) before anything is allocated. */
#define BRK_ALLOC 1         /* n = bytes                          -> value = ptr */
#define BRK_FREE 2          /* p0                                 (no reply) */
#define BRK_RAND 3          /* n = elements, s0 = seed            -> value = new f64 buffer */
#define BRK_EWISE 4         /* code = BCI_EW_*, p0 [, p1 | s0], n -> value = new buffer */
#define BRK_REDUCE 5        /* code = BCI_RED_*, pre, p0, n       -> value = result bits */
#define BRK_REDUCE_FUSED 6  /* + aux = ew op, p1 | s0, s1 addend  -> value = result bits */
#define BRK_ARGEXT 7        /* code = BCI_ARG*, p0, n             -> value = index */
#define BRK_CUMSUM 8        /* p0, n                              -> value = new buffer */
#define BRK_SORT 9          /* p0, n, flag INPLACE                -> value = buffer */
#define BRK_ARGSORT 10      /* p0, n                              -> value = new int64 buffer */
#define BRK_MATMUL 11       /* p0 = A, p1 = B, i0,i1,i2 = M,N,K   -> value = new buffer */
#define BRK_COPY 12         /* p0, n = bytes                      -> value = new buffer */
#define BRK_UPLOAD 13       /* n = bytes; payload = data | descriptor (flag FD), i0 = file offset -> value = new buffer */
#define BRK_DOWNLOAD 14     /* p0, i1 = byte offset into the buffer, n = bytes; descriptor (flag FD), i0 = file offset, or none -> payload = data */
#define BRK_SHA256_DEV 15   /* p0, n = bytes                      -> payload = 32-byte digest */
#define BRK_SHA256_FD 16    /* descriptor (flag FD) of a regular file -> payload = digest of its bytes */
#define BRK_SHA256_BYTES 17 /* payload = data                     -> payload = digest */
#define BRK_SHA256_UPDATE 18 /* payload = [state (sizeof bci_sha256_state) if flag STATE] + chunk -> payload = state */
#define BRK_SHA256_FINAL 19 /* payload = state                    -> payload = digest */
#define BRK_PING 20         /*                                    -> payload = 4 x u64 stats */
#define BRK_RAND_NORMAL 22  /* n = elements, i0 = seed, s0 = loc bits, s1 = scale bits -> value = new f64 buffer */
#define BRK_RAND_PCG 23     /* code = BCI_PCG_*, n = elements, i0 = low, s0 = range; payload = state lo,hi, inc lo,hi (4 x u64)
                               -> value = new buffer, payload = u32 rejected draws (bounded kinds) */
#define BRK_REDUCE_RAND 24  /* code = BCI_RED_*, pre, n = elements, s0 = seed -> value = result bits (float64) */
#define BRK_FILL 25         /* code = BCI_FILL_*, dtype, n, s0 = a, s1 = b [, i0 = c bits, flag ADDEND] -> value = new buffer */
#define BRK_CAST 26         /* dtype = source, aux = destination dtype, p0, n -> value = new buffer */
#define BRK_EXPORT 27       /* p0, n = bytes, descriptor (flag FD), i0 = file offset; payload = bytes that precede the
                               array in the file (an .npy header) -> the array is written behind them through the pinned
                               path while the device hashes header||array; payload = digest of the file as written */
#define BRK_RAND_MT 21      /* n = elements; payload = MT19937 state (625 x u32) -> value = new f64 buffer, payload = state after */

#define BRK_F_IN1 1u      /* p1 is an array operand (else scalar s0) */
#define BRK_F_ADDEND 2u   /* s1 is an addend (BRK_REDUCE_FUSED) */
#define BRK_F_INPLACE 4u
#define BRK_F_FD 8u       /* ONE file descriptor rides on this request as SCM_RIGHTS ancillary data: bulk bytes
                             move through it (a memfd or an unlinked tmpfs file the client made, or the output
                             file itself).  The broker never opens a path a client names: it can only touch
                             what the client could already open. */
#define BRK_F_STATE 16u   /* payload starts with a hash state */

#pragma pack(push, 1)
typedef struct bci_wire_req {
  uint32_t op, code, dtype, pre;
  uint32_t aux, flags;
  uint64_t p0, p1;
  uint64_t n;
  int64_t i0, i1, i2;
  uint64_t s0, s1;      /* bci_scalar bit patterns (double or int64) */
  uint32_t payload_len;
  uint32_t reserved;
} bci_wire_req;
typedef struct bci_wire_rep {
  int32_t status;
  uint32_t payload_len;
  uint64_t value;
} bci_wire_rep;
#pragma pack(pop)

#ifdef __cplusplus
}
#endif
#endif /* BCI_BROKER_H */
