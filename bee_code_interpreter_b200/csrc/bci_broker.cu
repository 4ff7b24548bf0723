// bci_broker.cu — native per-GPU broker (see include/bci_broker.h).
//
// Threads: one acceptor, one per connection (blocking reads of 96-byte requests), TWO
// dispatchers.  Connection threads validate buffer ownership, allocate result buffers and
// enqueue a Job; a dispatcher sleeps `coalesce_us` after the first arrival, drains the
// queue, groups jobs by op class and issues ONE batched call of the C ABI (bci.h) per group —
// the "one launch per op class for all concurrent requests" of the north star — then wakes
// the waiting connection threads with their results.  A failing group is retried job by job
// so one bad request cannot fail its neighbours.  Two dispatchers because a result-returning
// call (reduce, argmax) waits for the GPU outside the library's lock (bci_runtime.cu result
// slots): while one dispatcher waits, the other drains what arrived meanwhile and launches it
// behind the running work, so the op stream does not run dry between dispatch cycles.
#include "bci_common.cuh"
#include "../../include/bci_broker.h"

#include <sys/socket.h>
#include <sys/un.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <errno.h>
#include <signal.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace bci {
namespace {

struct Job {
  bci_wire_req rq;
  uint64_t out_ptr = 0;             // result buffer (already allocated by the connection thread)
  const void* host_data = nullptr;  // sha256 of host bytes
  size_t host_len = 0;
  bci_sha256_state* state = nullptr;
  uint32_t* mt_state = nullptr;     // BRK_RAND_MT: 625 words, updated in place
  bci_pcg64_gen pcg{};              // BRK_RAND_PCG
  uint32_t rejected = 0;
  // results
  int status = BCI_OK;
  uint64_t value = 0;
  uint8_t digest[32];
  std::string error;
  bool done = false;
};

using Key = std::tuple<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, int64_t, int64_t, int64_t>;

struct Broker {
  std::string path;
  int listen_fd = -1;
  int coalesce_us = 150;
  std::atomic<bool> running{false};
  static constexpr int kDispatchers = 2;
  std::thread acceptor, dispatcher[kDispatchers];
  std::mutex mu;
  std::condition_variable cv_jobs, cv_done;
  std::deque<Job*> queue;
  std::atomic<uint64_t> n_jobs{0}, n_batches{0}, max_batch{0}, n_conn{0};
  // connection workers: a fixed pool fed by the acceptor.  One script process = one connection, so
  // connections arrive at the request rate; creating a thread per connection inside a process
  // that holds a CUDA context (a very large address space) was the broker's per-request ceiling.
  static constexpr int kConnWorkers = 384;
  std::vector<std::thread> conn_workers;
  std::mutex conn_mu;
  std::condition_variable cv_conn;
  std::deque<int> conn_fds;
};

Broker g;

Key key_of(const bci_wire_req& r) {
  // everything that must be equal for two requests to share one launch
  switch (r.op) {
    case BRK_RAND: return Key{r.op, 0, 0, 0, 0, 0, 0, 0, 0};
    case BRK_EWISE: return Key{r.op, r.code, r.dtype, 0, 0, r.flags & BRK_F_IN1, 0, 0, 0};
    case BRK_REDUCE: return Key{r.op, r.code, r.dtype, r.pre, 0, 0, 0, 0, 0};
    case BRK_REDUCE_FUSED: return Key{r.op, r.code, r.dtype, r.pre, r.aux, r.flags & (BRK_F_IN1 | BRK_F_ADDEND), 0, 0, 0};
    case BRK_ARGEXT: return Key{r.op, r.code, r.dtype, 0, 0, 0, 0, 0, 0};
    case BRK_CUMSUM: case BRK_SORT: case BRK_ARGSORT: return Key{r.op, 0, r.dtype, 0, 0, 0, 0, 0, 0};
    case BRK_MATMUL: return Key{r.op, 0, 0, 0, 0, 0, r.i0, r.i1, r.i2};
    case BRK_RAND_PCG: return Key{r.op, r.code, 0, 0, 0, 0, 0, 0, 0};
    case BRK_REDUCE_RAND: return Key{r.op, r.code, 0, r.pre, 0, 0, 0, 0, 0};
    case BRK_FILL: return Key{r.op, r.code, r.dtype, 0, 0, r.flags & BRK_F_ADDEND, 0, 0, 0};
    case BRK_CAST: return Key{r.op, 0, r.dtype, 0, r.aux, 0, 0, 0, 0};
    default: return Key{r.op, 0, 0, 0, 0, 0, 0, 0, 0};
  }
}

bci_scalar scalar_of(uint64_t bits) {
  bci_scalar s;
  memcpy(&s, &bits, 8);
  return s;
}

// One batched C-ABI call for a group of same-class jobs.  Returns the call's status.
int run_group(std::vector<Job*>& grp) {
  const int B = (int)grp.size();
  const bci_wire_req& h = grp[0]->rq;
  std::vector<const void*> in0(B), in1(B);
  std::vector<void*> out(B);
  std::vector<size_t> n(B);
  for (int i = 0; i < B; ++i) {
    in0[i] = reinterpret_cast<const void*>(grp[i]->rq.p0);
    in1[i] = reinterpret_cast<const void*>(grp[i]->rq.p1);
    out[i] = reinterpret_cast<void*>(grp[i]->out_ptr);
    n[i] = (size_t)grp[i]->rq.n;
  }
  switch (h.op) {
    case BRK_RAND: {
      std::vector<uint64_t> seeds(B);
      for (int i = 0; i < B; ++i) seeds[i] = grp[i]->rq.s0;
      return bci_rand_f64(B, out.data(), n.data(), seeds.data());
    }
    case BRK_RAND_NORMAL: {
      std::vector<uint64_t> seeds(B);
      std::vector<double> loc(B), scale(B);
      for (int i = 0; i < B; ++i) {
        seeds[i] = (uint64_t)grp[i]->rq.i0;
        loc[i] = scalar_of(grp[i]->rq.s0).f;
        scale[i] = scalar_of(grp[i]->rq.s1).f;
      }
      return bci_rand_normal_f64(B, out.data(), n.data(), seeds.data(), loc.data(), scale.data());
    }
    case BRK_RAND_MT: {
      std::vector<uint32_t> st((size_t)B * BCI_MT19937_STATE_WORDS);
      for (int i = 0; i < B; ++i) memcpy(&st[(size_t)i * BCI_MT19937_STATE_WORDS], grp[i]->mt_state, 4 * BCI_MT19937_STATE_WORDS);
      int rc = bci_rand_mt19937_f64(B, out.data(), n.data(), st.data());
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) memcpy(grp[i]->mt_state, &st[(size_t)i * BCI_MT19937_STATE_WORDS], 4 * BCI_MT19937_STATE_WORDS);
      return rc;
    }
    case BRK_RAND_PCG: {
      std::vector<bci_pcg64_gen> gen(B);
      std::vector<uint32_t> rej(B, 0);
      for (int i = 0; i < B; ++i) gen[i] = grp[i]->pcg;
      const bool bounded = h.code == BCI_PCG_BOUNDED32_I64 || h.code == BCI_PCG_BOUNDED32_I32;
      int rc = bci_rand_pcg64((int)h.code, B, out.data(), n.data(), gen.data(), bounded ? rej.data() : nullptr);
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) grp[i]->rejected = rej[i];
      return rc;
    }
    case BRK_REDUCE_RAND: {
      std::vector<uint64_t> seeds(B), res(B, 0);
      for (int i = 0; i < B; ++i) seeds[i] = grp[i]->rq.s0;
      int rc = bci_reduce_rand_f64((int)h.code, (int)h.pre, B, n.data(), seeds.data(), res.data());
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) grp[i]->value = res[i];
      return rc;
    }
    case BRK_FILL: {
      std::vector<bci_scalar> a(B), b(B), c(B);
      for (int i = 0; i < B; ++i) { a[i] = scalar_of(grp[i]->rq.s0); b[i] = scalar_of(grp[i]->rq.s1); c[i] = scalar_of((uint64_t)grp[i]->rq.i0); }
      return bci_fill((int)h.code, (int)h.dtype, B, out.data(), n.data(), a.data(), b.data(), (h.flags & BRK_F_ADDEND) ? c.data() : nullptr);
    }
    case BRK_CAST: return bci_cast((int)h.dtype, (int)h.aux, B, in0.data(), out.data(), n.data());
    case BRK_EWISE: {
      std::vector<bci_scalar> sc(B);
      for (int i = 0; i < B; ++i) sc[i] = scalar_of(grp[i]->rq.s0);
      const bool arr = h.flags & BRK_F_IN1;
      return bci_ewise((int)h.code, (int)h.dtype, B, in0.data(), arr ? in1.data() : nullptr, sc.data(), out.data(), n.data());
    }
    case BRK_REDUCE: case BRK_REDUCE_FUSED: {
      std::vector<uint64_t> res(B, 0);
      int rc;
      if (h.op == BRK_REDUCE) {
        rc = bci_reduce((int)h.code, (int)h.pre, (int)h.dtype, B, in0.data(), n.data(), res.data());
      } else {
        std::vector<bci_scalar> sc(B), ad(B);
        for (int i = 0; i < B; ++i) { sc[i] = scalar_of(grp[i]->rq.s0); ad[i] = scalar_of(grp[i]->rq.s1); }
        rc = bci_reduce_fused((int)h.code, (int)h.pre, (int)h.aux, (int)h.dtype, B, in0.data(),
                              (h.flags & BRK_F_IN1) ? in1.data() : nullptr, sc.data(),
                              (h.flags & BRK_F_ADDEND) ? ad.data() : nullptr, n.data(), res.data());
      }
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) grp[i]->value = res[i];
      return rc;
    }
    case BRK_ARGEXT: {
      std::vector<int64_t> res(B, 0);
      int rc = bci_argext((int)h.code, (int)h.dtype, B, in0.data(), n.data(), res.data());
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) grp[i]->value = (uint64_t)res[i];
      return rc;
    }
    case BRK_CUMSUM: return bci_cumsum((int)h.dtype, B, in0.data(), out.data(), n.data());
    case BRK_SORT: return bci_sort((int)h.dtype, B, in0.data(), out.data(), n.data());
    case BRK_ARGSORT: return bci_argsort((int)h.dtype, B, in0.data(), reinterpret_cast<int64_t* const*>(out.data()), n.data());
    case BRK_MATMUL: return bci_matmul_f32(B, in0.data(), in1.data(), out.data(), (int)h.i0, (int)h.i1, (int)h.i2);
    case BRK_SHA256_DEV: case BRK_SHA256_BYTES: {
      std::vector<uint8_t> dig((size_t)B * 32);
      const bool on_dev = h.op == BRK_SHA256_DEV;
      if (!on_dev) for (int i = 0; i < B; ++i) { in0[i] = grp[i]->host_data; n[i] = grp[i]->host_len; }
      int rc = bci_sha256(B, in0.data(), n.data(), on_dev ? 1 : 0, dig.data());
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) memcpy(grp[i]->digest, &dig[(size_t)i * 32], 32);
      return rc;
    }
    case BRK_SHA256_UPDATE: {
      std::vector<bci_sha256_state*> st(B);
      for (int i = 0; i < B; ++i) { st[i] = grp[i]->state; in0[i] = grp[i]->host_data; n[i] = grp[i]->host_len; }
      return bci_sha256_update_batch(B, st.data(), in0.data(), n.data());
    }
    case BRK_SHA256_FINAL: {
      std::vector<bci_sha256_state*> st(B);
      std::vector<uint8_t> dig((size_t)B * 32);
      for (int i = 0; i < B; ++i) st[i] = grp[i]->state;
      int rc = bci_sha256_final_batch(B, st.data(), dig.data());
      if (rc == BCI_OK) for (int i = 0; i < B; ++i) memcpy(grp[i]->digest, &dig[(size_t)i * 32], 32);
      return rc;
    }
  }
  set_error("broker: op %u cannot be batched", h.op);
  return BCI_ERR_INVALID;
}

// Runs one group and wakes its waiters.  A failing group is retried job by job so that one bad request
// cannot fail its neighbours.
void complete_group(std::vector<Job*>& grp) {
  int rc = run_group(grp);
  if (rc != BCI_OK && grp.size() > 1) {
    for (Job* j : grp) {  // isolate the offender
      std::vector<Job*> one{j};
      j->status = run_group(one);
      if (j->status != BCI_OK) j->error = bci_last_error();
    }
  } else {
    const std::string err = rc != BCI_OK ? bci_last_error() : "";
    for (Job* j : grp) { j->status = rc; j->error = err; }
  }
  {
    std::lock_guard<std::mutex> lk(g.mu);
    for (Job* j : grp) j->done = true;
  }
  g.cv_done.notify_all();
}

void dispatch_loop() {
  while (g.running.load()) {
    std::vector<Job*> jobs;
    {
      std::unique_lock<std::mutex> lk(g.mu);
      g.cv_jobs.wait(lk, [] { return !g.queue.empty() || !g.running.load(); });
      if (!g.running.load() && g.queue.empty()) break;
    }
    if (g.coalesce_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(g.coalesce_us));
    {
      std::lock_guard<std::mutex> lk(g.mu);
      jobs.assign(g.queue.begin(), g.queue.end());
      g.queue.clear();
    }
    std::map<Key, std::vector<Job*>> groups;
    for (Job* j : jobs) groups[key_of(j->rq)].push_back(j);
    for (auto& kv : groups) {
      std::vector<Job*>& grp = kv.second;
      g.n_batches++;
      g.n_jobs += grp.size();
      if (grp.size() > g.max_batch.load()) g.max_batch = grp.size();
      const uint32_t op = grp[0]->rq.op;
      const bool hashing = op == BRK_SHA256_DEV || op == BRK_SHA256_BYTES || op == BRK_SHA256_UPDATE || op == BRK_SHA256_FINAL;
      if (hashing) {
        // a hashing launch lasts as long as its longest message (one SHA-256 is a serial chain: ~1 s for
        // 64 MiB) and runs on a side stream: it gets its own thread so that the dispatchers keep packing
        // everybody else's ops meanwhile
        std::thread([grp]() mutable { complete_group(grp); }).detach();
      } else {
        complete_group(grp);
      }
    }
  }
}

void submit_and_wait(Job& job) {
  std::unique_lock<std::mutex> lk(g.mu);
  g.queue.push_back(&job);
  g.cv_jobs.notify_one();
  g.cv_done.wait(lk, [&] { return job.done; });
}

// Reads exactly n bytes; file descriptors that arrive as SCM_RIGHTS ancillary data on the way are
// queued on the connection (a request with BRK_F_FD consumes one; whatever a client sends beyond
// that is closed at the end of the request).
bool read_full(int fd, void* buf, size_t n, std::deque<int>* fds) {
  uint8_t* p = reinterpret_cast<uint8_t*>(buf);
  while (n) {
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int) * 8)];
    iovec iov{p, n};
    msghdr mh{};
    mh.msg_iov = &iov; mh.msg_iovlen = 1;
    mh.msg_control = ctl; mh.msg_controllen = sizeof ctl;
    ssize_t r = ::recvmsg(fd, &mh, MSG_CMSG_CLOEXEC);
    if (r == 0) return false;
    if (r < 0) { if (errno == EINTR) continue; return false; }
    for (cmsghdr* cm = CMSG_FIRSTHDR(&mh); cm; cm = CMSG_NXTHDR(&mh, cm)) {
      if (cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) continue;
      const size_t cnt = (cm->cmsg_len - CMSG_LEN(0)) / sizeof(int);
      for (size_t i = 0; i < cnt; ++i) {
        int got;
        memcpy(&got, CMSG_DATA(cm) + i * sizeof(int), sizeof(int));
        if (fds && fds->size() < 8) fds->push_back(got); else ::close(got);
      }
    }
    p += r; n -= (size_t)r;
  }
  return true;
}
bool write_full(int fd, const void* buf, size_t n) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(buf);
  while (n) {
    ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
    if (r < 0) { if (errno == EINTR) continue; return false; }
    p += r; n -= (size_t)r;
  }
  return true;
}
bool reply(int fd, int status, uint64_t value, const void* payload, size_t len) {
  bci_wire_rep rp;
  rp.status = status; rp.payload_len = (uint32_t)len; rp.value = value;
  if (!write_full(fd, &rp, sizeof rp)) return false;
  return len == 0 || write_full(fd, payload, len);
}
bool reply_err(int fd, int status, const std::string& msg) { return reply(fd, status, 0, msg.data(), msg.size()); }

struct FdGuard {
  int fd = -1;
  ~FdGuard() { if (fd >= 0) ::close(fd); }
};

// n * es without wrapping (sizes come from untrusted scripts)
bool mul_ok(uint64_t a, uint64_t b, size_t* out) {
  if (b != 0 && a > (uint64_t)SIZE_MAX / b) return false;
  *out = (size_t)(a * b);
  return true;
}

// Digests of files this broker wrote from device memory (BRK_EXPORT), keyed by file identity: a later
// BRK_SHA256_FD of the same untouched file is answered without reading it back.  Any write()/truncate
// by the script changes mtime/ctime and invalidates the entry.  (A writable shared mapping of a tmpfs
// file does not: deployments that do not trust scripts start the broker with export digests off.)
struct ExportKey { uint64_t dev, ino; bool operator<(const ExportKey& o) const { return dev != o.dev ? dev < o.dev : ino < o.ino; } };
struct ExportVal { uint64_t size; int64_t mtime_ns, ctime_ns; uint8_t digest[32]; };
std::mutex g_export_mu;
std::map<ExportKey, ExportVal> g_exports;
bool g_trust_exports = true;

void export_remember(const struct stat& st, const uint8_t* digest) {
  if (!g_trust_exports) return;
  ExportVal v;
  v.size = (uint64_t)st.st_size;
  v.mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
  v.ctime_ns = (int64_t)st.st_ctim.tv_sec * 1000000000ll + st.st_ctim.tv_nsec;
  memcpy(v.digest, digest, 32);
  std::lock_guard<std::mutex> lk(g_export_mu);
  if (g_exports.size() > 16384) g_exports.clear();
  g_exports[ExportKey{(uint64_t)st.st_dev, (uint64_t)st.st_ino}] = v;
}
bool export_lookup(const struct stat& st, uint8_t* digest) {
  std::lock_guard<std::mutex> lk(g_export_mu);
  auto it = g_exports.find(ExportKey{(uint64_t)st.st_dev, (uint64_t)st.st_ino});
  if (it == g_exports.end()) return false;
  const ExportVal& v = it->second;
  const int64_t m = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
  const int64_t c = (int64_t)st.st_ctim.tv_sec * 1000000000ll + st.st_ctim.tv_nsec;
  if (v.size != (uint64_t)st.st_size || v.mtime_ns != m || v.ctime_ns != c) { g_exports.erase(it); return false; }
  memcpy(digest, v.digest, 32);
  g_exports.erase(it);      // one consumer: the executor stores the file right after
  return true;
}

size_t out_elem_size(uint32_t op, uint32_t dtype) {
  if (op == BRK_ARGSORT) return 8;
  if (op == BRK_CUMSUM && dtype == BCI_I32) return 8;
  return elem_size((int)dtype);
}

void serve(int fd) {
  std::unordered_map<uint64_t, size_t> owned;
  std::deque<int> fds;          // descriptors received with the current request
  auto owns = [&](uint64_t p, size_t bytes) {
    auto it = owned.find(p);
    return it != owned.end() && bytes <= it->second;
  };
  auto new_buf = [&](size_t bytes, uint64_t* out) {
    void* d = nullptr;
    int rc = bci_alloc(bytes, &d);
    if (rc == BCI_OK) { *out = reinterpret_cast<uint64_t>(d); owned[*out] = bytes ? bytes : 1; }
    return rc;
  };
  auto drop = [&](uint64_t p) {
    if (owned.erase(p)) bci_free(reinterpret_cast<void*>(p));
  };
  auto take_fd = [&](FdGuard& g) {   // the ONE descriptor a BRK_F_FD request carries
    if (fds.empty()) return false;
    g.fd = fds.front();
    fds.pop_front();
    return true;
  };
  std::vector<uint8_t> payload;
  for (;;) {
    bci_wire_req rq;
    if (!read_full(fd, &rq, sizeof rq, &fds)) break;
    if (rq.payload_len > (64u << 20)) break;                       // no request body is anywhere near this
    payload.resize(rq.payload_len);
    if (rq.payload_len && !read_full(fd, payload.data(), rq.payload_len, &fds)) break;
    const size_t es = elem_size((int)rq.dtype);
    bool ok = true;
    switch (rq.op) {
      case BRK_FREE: drop(rq.p0); break;
      case BRK_ALLOC: {
        uint64_t p = 0;
        int rc = new_buf((size_t)rq.n, &p);
        ok = rc == BCI_OK ? reply(fd, BCI_OK, p, nullptr, 0) : reply_err(fd, rc, bci_last_error());
        break;
      }
      case BRK_PING: {
        uint64_t st[4] = {g.n_jobs.load(), g.n_batches.load(), g.max_batch.load(), g.n_conn.load()};
        ok = reply(fd, BCI_OK, 0, st, sizeof st);
        break;
      }
      case BRK_COPY: {
        if (!owns(rq.p0, (size_t)rq.n)) { ok = reply_err(fd, BCI_ERR_INVALID, "buffer does not belong to this request"); break; }
        uint64_t p = 0;
        int rc = new_buf((size_t)rq.n, &p);
        if (rc == BCI_OK) rc = bci_d2d(reinterpret_cast<void*>(p), reinterpret_cast<void*>(rq.p0), (size_t)rq.n);
        ok = rc == BCI_OK ? reply(fd, BCI_OK, p, nullptr, 0) : reply_err(fd, rc, bci_last_error());
        break;
      }
      case BRK_UPLOAD: {
        FdGuard src;
        if ((rq.flags & BRK_F_FD) && !take_fd(src)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: request announces a descriptor but none arrived"); break; }
        uint64_t p = 0;
        int rc = new_buf((size_t)rq.n, &p);
        if (rc == BCI_OK && rq.n) {
          if (rq.flags & BRK_F_FD) {
            rc = rq.i0 < 0 ? BCI_ERR_INVALID : bci_h2d_fd(reinterpret_cast<void*>(p), src.fd, (uint64_t)rq.i0, (size_t)rq.n);
          } else if (payload.size() >= rq.n) {
            rc = bci_h2d(reinterpret_cast<void*>(p), payload.data(), (size_t)rq.n);
            if (rc == BCI_OK) rc = bci_sync();
          } else { rc = BCI_ERR_INVALID; set_error("broker: short upload payload"); }
        }
        if (rc != BCI_OK && p) drop(p);
        ok = rc == BCI_OK ? reply(fd, BCI_OK, p, nullptr, 0) : reply_err(fd, rc, bci_last_error());
        break;
      }
      case BRK_DOWNLOAD: {
        FdGuard dst;
        if ((rq.flags & BRK_F_FD) && !take_fd(dst)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: request announces a descriptor but none arrived"); break; }
        // i1 = byte offset INTO the buffer: `x[i]`, `x[a:b]` of a device array fetch only what they print
        size_t end = 0;
        if (rq.i1 < 0 || __builtin_add_overflow((size_t)rq.i1, (size_t)rq.n, &end) || !owns(rq.p0, end)) {
          ok = reply_err(fd, BCI_ERR_INVALID, "buffer does not belong to this request"); break;
        }
        const uint8_t* src = reinterpret_cast<const uint8_t*>(rq.p0) + rq.i1;
        if (rq.flags & BRK_F_FD) {
          int rc = rq.i0 < 0 ? BCI_ERR_INVALID : bci_d2h_fd(dst.fd, (uint64_t)rq.i0, src, (size_t)rq.n);
          ok = rc == BCI_OK ? reply(fd, BCI_OK, 0, nullptr, 0) : reply_err(fd, rc, bci_last_error());
        } else if (rq.n > (64u << 20)) {
          ok = reply_err(fd, BCI_ERR_INVALID, "broker: inline downloads stop at 64 MiB, pass a descriptor");
        } else {
          std::vector<uint8_t> host((size_t)rq.n);
          int rc = rq.n ? bci_d2h(host.data(), src, (size_t)rq.n) : BCI_OK;
          ok = rc == BCI_OK ? reply(fd, BCI_OK, 0, host.data(), host.size()) : reply_err(fd, rc, bci_last_error());
        }
        break;
      }
      case BRK_EXPORT: {
        // array -> file (pinned path) while the device hashes header||array: the output file's id is
        // known before its bytes have left the GPU
        FdGuard dst;
        if (!(rq.flags & BRK_F_FD) || !take_fd(dst)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: export needs a descriptor"); break; }
        const size_t hdr = payload.size();
        if (!owns(rq.p0, (size_t)rq.n) || rq.i0 < 0) { ok = reply_err(fd, BCI_ERR_INVALID, "buffer does not belong to this request"); break; }
        struct stat st;
        if (fstat(dst.fd, &st) != 0 || !S_ISREG(st.st_mode)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: export target is not a regular file"); break; }
        size_t total = 0;
        if (__builtin_add_overflow(hdr, (size_t)rq.n, &total)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: size overflow"); break; }
        uint64_t msg = 0;
        int rc = new_buf(total, &msg);
        if (rc == BCI_OK && hdr) { rc = bci_h2d(reinterpret_cast<void*>(msg), payload.data(), hdr); if (rc == BCI_OK) rc = bci_sync(); }
        if (rc == BCI_OK && rq.n) rc = bci_d2d(reinterpret_cast<uint8_t*>(msg) + hdr, reinterpret_cast<void*>(rq.p0), (size_t)rq.n);
        if (rc != BCI_OK) { if (msg) drop(msg); ok = reply_err(fd, rc, bci_last_error()); break; }
        Job job;
        job.rq = rq;
        job.rq.op = BRK_SHA256_DEV;
        job.rq.p0 = msg;
        job.rq.n = total;
        // the copy-out runs on a transfer lane while the hash job waits in the batching queue
        std::string werr;      // bci_last_error() is thread-local: carry the writer's message over
        std::thread writer([&] {
          rc = bci_d2h_fd(dst.fd, (uint64_t)rq.i0, reinterpret_cast<void*>(msg), total);
          if (rc != BCI_OK) werr = bci_last_error();
        });
        submit_and_wait(job);
        writer.join();
        drop(msg);
        if (rc != BCI_OK) set_error("%s", werr.c_str());
        else if (job.status != BCI_OK) { rc = job.status; set_error("%s", job.error.c_str()); }
        if (rc == BCI_OK && rq.i0 == 0) {
          if (ftruncate(dst.fd, (off_t)total) == 0 && fstat(dst.fd, &st) == 0) export_remember(st, job.digest);
        }
        ok = rc == BCI_OK ? reply(fd, BCI_OK, 0, job.digest, 32) : reply_err(fd, rc, bci_last_error());
        break;
      }
      case BRK_RAND: case BRK_RAND_NORMAL: case BRK_RAND_PCG: case BRK_REDUCE_RAND: case BRK_FILL: case BRK_CAST:
      case BRK_EWISE: case BRK_CUMSUM: case BRK_SORT: case BRK_ARGSORT: case BRK_MATMUL:
      case BRK_REDUCE: case BRK_REDUCE_FUSED: case BRK_ARGEXT: case BRK_SHA256_DEV: {
        Job job;
        job.rq = rq;
        // ownership + sizes (every product is checked for overflow: the numbers come from scripts)
        bool valid = true;
        size_t in_bytes = 0, out_bytes = 0;
        if (rq.op == BRK_RAND || rq.op == BRK_RAND_NORMAL) { valid = mul_ok(rq.n, 8, &out_bytes); }
        else if (rq.op == BRK_REDUCE_RAND) { valid = rq.n < (1ull << 48); }
        else if (rq.op == BRK_RAND_PCG) {
          const size_t oes = rq.code == BCI_PCG_F64 || rq.code == BCI_PCG_BOUNDED32_I64 ? 8 : 4;
          valid = rq.code <= BCI_PCG_BOUNDED32_I32 && payload.size() == 32 && mul_ok(rq.n, oes, &out_bytes);
          if (valid) {
            uint64_t w[4];
            memcpy(w, payload.data(), 32);
            job.pcg.state_lo = w[0]; job.pcg.state_hi = w[1]; job.pcg.inc_lo = w[2]; job.pcg.inc_hi = w[3];
            job.pcg.low = rq.i0; job.pcg.range = rq.s0;
          }
        }
        else if (rq.op == BRK_FILL) { valid = es != 0 && mul_ok(rq.n, es, &out_bytes); }
        else if (rq.op == BRK_CAST) {
          const size_t des = elem_size((int)rq.aux);
          valid = es != 0 && des != 0 && mul_ok(rq.n, es, &in_bytes) && mul_ok(rq.n, des, &out_bytes) && owns(rq.p0, in_bytes);
        }
        else if (rq.op == BRK_SHA256_DEV) { in_bytes = (size_t)rq.n; valid = owns(rq.p0, in_bytes); }
        else if (rq.op == BRK_MATMUL) {
          size_t a_b = 0, b_b = 0, t = 0;
          valid = rq.i0 > 0 && rq.i1 > 0 && rq.i2 > 0 && rq.i0 < (1ll << 31) && rq.i1 < (1ll << 31) && rq.i2 < (1ll << 31) &&
                  mul_ok((uint64_t)rq.i0, (uint64_t)rq.i2, &t) && mul_ok(t, 4, &a_b) &&
                  mul_ok((uint64_t)rq.i2, (uint64_t)rq.i1, &t) && mul_ok(t, 4, &b_b) &&
                  mul_ok((uint64_t)rq.i0, (uint64_t)rq.i1, &t) && mul_ok(t, 4, &out_bytes) &&
                  owns(rq.p0, a_b) && owns(rq.p1, b_b);
        } else {
          valid = es != 0 && mul_ok(rq.n, es, &in_bytes) && owns(rq.p0, in_bytes);
          if (valid && (rq.flags & BRK_F_IN1) && (rq.op == BRK_EWISE || rq.op == BRK_REDUCE_FUSED)) valid = owns(rq.p1, in_bytes);
          if (valid && (rq.op == BRK_EWISE || rq.op == BRK_CUMSUM || rq.op == BRK_SORT || rq.op == BRK_ARGSORT))
            valid = mul_ok(rq.n, out_elem_size(rq.op, rq.dtype), &out_bytes);
        }
        if (!valid) { ok = reply_err(fd, BCI_ERR_INVALID, "buffer does not belong to this request (or its size overflows)"); break; }
        const bool inplace = rq.op == BRK_SORT && (rq.flags & BRK_F_INPLACE);
        const bool needs_out = !inplace && (rq.op == BRK_RAND || rq.op == BRK_RAND_NORMAL || rq.op == BRK_RAND_PCG || rq.op == BRK_FILL ||
                                            rq.op == BRK_CAST || rq.op == BRK_EWISE || rq.op == BRK_CUMSUM ||
                                            rq.op == BRK_SORT || rq.op == BRK_ARGSORT || rq.op == BRK_MATMUL);
        if (inplace) job.out_ptr = rq.p0;
        if (needs_out) {
          int rc = new_buf(out_bytes, &job.out_ptr);
          if (rc != BCI_OK) { ok = reply_err(fd, rc, bci_last_error()); break; }
        }
        submit_and_wait(job);
        if (job.status != BCI_OK) {
          if (job.out_ptr && job.out_ptr != rq.p0) drop(job.out_ptr);
          ok = reply_err(fd, job.status, job.error);
        } else if (rq.op == BRK_SHA256_DEV) {
          ok = reply(fd, BCI_OK, 0, job.digest, 32);
        } else if (rq.op == BRK_RAND_PCG) {
          ok = reply(fd, BCI_OK, job.out_ptr, &job.rejected, 4);
        } else {
          const bool scalar_result = rq.op == BRK_REDUCE || rq.op == BRK_REDUCE_FUSED || rq.op == BRK_ARGEXT || rq.op == BRK_REDUCE_RAND;
          ok = reply(fd, BCI_OK, scalar_result ? job.value : job.out_ptr, nullptr, 0);
        }
        break;
      }
      case BRK_RAND_MT: {
        size_t out_bytes = 0;
        if (payload.size() != 4 * BCI_MT19937_STATE_WORDS || !mul_ok(rq.n, 8, &out_bytes)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: MT19937 state must be 625 words"); break; }
        Job job;
        job.rq = rq;
        uint32_t st[BCI_MT19937_STATE_WORDS];
        memcpy(st, payload.data(), sizeof st);
        job.mt_state = st;
        int rc = new_buf(out_bytes, &job.out_ptr);
        if (rc != BCI_OK) { ok = reply_err(fd, rc, bci_last_error()); break; }
        submit_and_wait(job);
        if (job.status != BCI_OK) { drop(job.out_ptr); ok = reply_err(fd, job.status, job.error); }
        else ok = reply(fd, BCI_OK, job.out_ptr, st, sizeof st);
        break;
      }
      case BRK_SHA256_FD: {
        // a file of the object store's caller: read through the pinned path into device memory and
        // hashed there with every other pending message; answered from the export table when this
        // broker wrote the file itself and nothing has touched it since
        FdGuard src;
        if (!(rq.flags & BRK_F_FD) || !take_fd(src)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: hashing a file needs its descriptor"); break; }
        struct stat st;
        if (fstat(src.fd, &st) != 0 || !S_ISREG(st.st_mode)) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: not a regular file"); break; }
        uint8_t digest[32];
        if (export_lookup(st, digest)) { ok = reply(fd, BCI_OK, 1, digest, 32); break; }
        const size_t len = (size_t)st.st_size;
        uint64_t msg = 0;
        int rc = new_buf(len, &msg);
        if (rc == BCI_OK && len) rc = bci_h2d_fd(reinterpret_cast<void*>(msg), src.fd, 0, len);
        if (rc != BCI_OK) { if (msg) drop(msg); ok = reply_err(fd, rc, bci_last_error()); break; }
        Job job;
        job.rq = rq;
        job.rq.op = BRK_SHA256_DEV;
        job.rq.p0 = msg;
        job.rq.n = len;
        submit_and_wait(job);
        drop(msg);
        ok = job.status == BCI_OK ? reply(fd, BCI_OK, 0, job.digest, 32) : reply_err(fd, job.status, job.error);
        break;
      }
      case BRK_SHA256_BYTES: {
        Job job;
        job.rq = rq;
        job.host_data = payload.data(); job.host_len = payload.size();
        submit_and_wait(job);
        ok = job.status == BCI_OK ? reply(fd, BCI_OK, 0, job.digest, 32) : reply_err(fd, job.status, job.error);
        break;
      }
      case BRK_SHA256_UPDATE: case BRK_SHA256_FINAL: {
        Job job;
        job.rq = rq;
        bci_sha256_state st;
        size_t off = 0;
        if (rq.op == BRK_SHA256_FINAL || (rq.flags & BRK_F_STATE)) {
          if (payload.size() < sizeof st) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: short hash state"); break; }
          memcpy(&st, payload.data(), sizeof st);
          if (st.buf_len >= 64) { ok = reply_err(fd, BCI_ERR_INVALID, "broker: corrupt hash state"); break; }
          off = sizeof st;
        } else {
          bci_sha256_init(&st);
        }
        job.state = &st;
        job.host_data = payload.data() + off;
        job.host_len = payload.size() - off;
        submit_and_wait(job);
        if (job.status != BCI_OK) ok = reply_err(fd, job.status, job.error);
        else if (rq.op == BRK_SHA256_FINAL) ok = reply(fd, BCI_OK, 0, job.digest, 32);
        else ok = reply(fd, BCI_OK, 0, &st, sizeof st);
        break;
      }
      default:
        ok = reply_err(fd, BCI_ERR_INVALID, "broker: unknown op");
    }
    for (int extra : fds) ::close(extra);      // descriptors the request did not use
    fds.clear();
    if (!ok) break;
  }
  for (int extra : fds) ::close(extra);
  for (auto& kv : owned) bci_free(reinterpret_cast<void*>(kv.first));
  ::close(fd);
}

void accept_loop() {
  while (g.running.load()) {
    int fd = ::accept(g.listen_fd, nullptr, nullptr);
    if (fd < 0) {
      if (errno == EINTR) continue;
      break;
    }
    g.n_conn++;
    {
      std::lock_guard<std::mutex> lk(g.conn_mu);
      g.conn_fds.push_back(fd);
    }
    g.cv_conn.notify_one();
  }
}

void conn_worker_loop() {
  for (;;) {
    int fd = -1;
    {
      std::unique_lock<std::mutex> lk(g.conn_mu);
      g.cv_conn.wait(lk, [] { return !g.conn_fds.empty() || !g.running.load(); });
      if (g.conn_fds.empty()) return;  // stopping
      fd = g.conn_fds.front();
      g.conn_fds.pop_front();
    }
    serve(fd);
  }
}

}  // namespace
}  // namespace bci

using namespace bci;

extern "C" int bci_broker_start(const char* socket_path, int coalesce_us) {
  BCI_REQUIRE_INIT();
  if (!socket_path) { set_error("bci_broker_start: NULL path"); return BCI_ERR_INVALID; }
  if (g.running.load()) { set_error("bci_broker_start: already running"); return BCI_ERR_INVALID; }
  sockaddr_un addr;
  memset(&addr, 0, sizeof addr);
  addr.sun_family = AF_UNIX;
  if (strlen(socket_path) >= sizeof addr.sun_path) { set_error("bci_broker_start: socket path too long"); return BCI_ERR_INVALID; }
  strcpy(addr.sun_path, socket_path);
  ::unlink(socket_path);
  int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
  if (fd < 0 || ::bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0 || ::listen(fd, 4096) != 0) {
    set_error("bci_broker_start: cannot listen on %s: %s", socket_path, strerror(errno));
    if (fd >= 0) ::close(fd);
    return BCI_ERR_INVALID;
  }
  {
    const char* t = getenv("BCI_TRUST_EXPORT_DIGESTS");   // "0": always re-read files the script could have touched
    g_trust_exports = !(t && t[0] == '0');
  }
  g.path = socket_path;
  g.listen_fd = fd;
  g.coalesce_us = coalesce_us;
  g.running = true;
  for (auto& d : g.dispatcher) d = std::thread(dispatch_loop);
  g.conn_workers.clear();
  for (int i = 0; i < Broker::kConnWorkers; ++i) g.conn_workers.emplace_back(conn_worker_loop);
  g.acceptor = std::thread(accept_loop);
  return BCI_OK;
}

extern "C" void bci_broker_stop(void) {
  if (!g.running.exchange(false)) return;
  ::shutdown(g.listen_fd, SHUT_RDWR);
  ::close(g.listen_fd);
  g.cv_jobs.notify_all();
  if (g.acceptor.joinable()) g.acceptor.join();
  for (auto& d : g.dispatcher)
    if (d.joinable()) d.join();
  {
    // wake idle workers; a worker still serving a connection returns when its client goes away
    std::lock_guard<std::mutex> lk(g.conn_mu);
    for (int fd : g.conn_fds) ::close(fd);
    g.conn_fds.clear();
  }
  g.cv_conn.notify_all();
  for (auto& w : g.conn_workers) {
    if (!w.joinable()) continue;
    w.detach();   // never block shutdown on a client that keeps its socket open
  }
  g.conn_workers.clear();
  ::unlink(g.path.c_str());
}

extern "C" int bci_broker_stats(uint64_t stats[4]) {
  if (!stats) return BCI_ERR_INVALID;
  stats[0] = g.n_jobs.load(); stats[1] = g.n_batches.load(); stats[2] = g.max_batch.load(); stats[3] = g.n_conn.load();
  return BCI_OK;
}
