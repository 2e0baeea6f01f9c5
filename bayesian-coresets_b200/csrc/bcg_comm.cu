// bcg_comm_*: a torch-free process group for N-sharding over the GPUs of ONE node (include/bcg.h).
//
// What it is for: bootstrap (row counts, the 64-byte cudaIpcMemHandle_t of every rank's mailbox) and the one-off /
// per-pass reductions of S-vectors (b, sum ||a_n||, the column sums of a SparseVI / BatchPSVI projection pass).  These
// are <= 8 KB messages a few times per job; the per-iteration exchange of the greedy loop never comes here -- it is
// fused into the kernels over NVLink peer memory (loop_kernel.cuh, step_kernels.cuh).  Hence plain TCP on the loopback
// interface, star topology: rank 0 listens, ranks 1..W-1 connect once.  An all-gather is "everyone sends to rank 0,
// rank 0 sends the concatenation back"; an all-reduce is an all-gather followed by the SAME rank-ordered reduction on
// every rank, so all ranks hold bit-identical results (the solvers rely on replicated float64 state).
// Host-only code (no CUDA calls): usable on a box without a GPU, which is how tests/test_native_comm.py runs it.
#include <arpa/inet.h>
#include <errno.h>
#include <math.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#include <vector>

#include "../../include/bcg.h"

extern "C" const char* bcg_last_error(void);
int bcg_set_error(int code, const char* fmt, ...);   // bcg_api.cu

struct bcg_comm {
  int rank, world;
  int timeout_ms;
  int listen_fd;
  std::vector<int> fds;          // rank 0: fds[r] = socket of rank r (fds[0] unused); others: fds[0] = socket to rank 0
  std::vector<unsigned char> buf;
};

namespace {
const uint32_t kMagic = 0x62636731u;   // "bcg1"

double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

bool wait_fd(int fd, short ev, int timeout_ms) {
  pollfd p{fd, ev, 0};
  for (;;) {
    const int r = poll(&p, 1, timeout_ms);
    if (r > 0) return true;
    if (r == 0) return false;
    if (errno != EINTR) return false;
  }
}

bool send_all(int fd, const void* data, size_t n, int timeout_ms) {
  const char* p = static_cast<const char*>(data);
  while (n > 0) {
    if (!wait_fd(fd, POLLOUT, timeout_ms)) return false;
    const ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) { if (errno == EINTR || errno == EAGAIN) continue; return false; }
    p += k; n -= (size_t)k;
  }
  return true;
}

bool recv_all(int fd, void* data, size_t n, int timeout_ms) {
  char* p = static_cast<char*>(data);
  while (n > 0) {
    if (!wait_fd(fd, POLLIN, timeout_ms)) return false;
    const ssize_t k = recv(fd, p, n, 0);
    if (k == 0) return false;
    if (k < 0) { if (errno == EINTR || errno == EAGAIN) continue; return false; }
    p += k; n -= (size_t)k;
  }
  return true;
}

void tune(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}
}  // namespace

extern "C" int bcg_comm_create(const char* addr, int32_t port, int32_t rank, int32_t world, int32_t timeout_ms,
                               bcg_comm** out) {
  if (!out) return bcg_set_error(BCG_ERR_ARG, "null out");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world || port <= 0 || port > 65535)
    return bcg_set_error(BCG_ERR_ARG, "bad rank/world/port %d/%d/%d", rank, world, port);
  if (timeout_ms <= 0) timeout_ms = 120000;
  bcg_comm* c = new bcg_comm();
  c->rank = rank; c->world = world; c->timeout_ms = timeout_ms; c->listen_fd = -1;
  c->fds.assign(world, -1);
  if (world == 1) { *out = c; return BCG_OK; }
  sockaddr_in sa;
  memset(&sa, 0, sizeof(sa));
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, (addr && *addr) ? addr : "127.0.0.1", &sa.sin_addr) != 1) {
    delete c;
    return bcg_set_error(BCG_ERR_ARG, "bad IPv4 address '%s'", addr ? addr : "");
  }
  const double t_end = now_ms() + timeout_ms;
  if (rank == 0) {
    const int lf = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(lf, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (lf < 0 || bind(lf, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0 || listen(lf, world) != 0) {
      const int e = errno;
      if (lf >= 0) close(lf);
      delete c;
      return bcg_set_error(BCG_ERR_COMM, "rank 0 cannot listen on %s:%d: %s", addr ? addr : "127.0.0.1", port, strerror(e));
    }
    c->listen_fd = lf;
    int have = 0;
    while (have < world - 1) {
      const int left = (int)(t_end - now_ms());
      if (left <= 0 || !wait_fd(lf, POLLIN, left)) {
        bcg_comm_destroy(c);
        return bcg_set_error(BCG_ERR_COMM, "rank 0: only %d of %d peers connected within %d ms", have, world - 1, timeout_ms);
      }
      const int fd = accept(lf, nullptr, nullptr);
      if (fd < 0) continue;
      tune(fd);
      uint32_t hello[3] = {0, 0, 0};
      if (!recv_all(fd, hello, sizeof(hello), 5000) || hello[0] != kMagic || hello[2] != (uint32_t)world ||
          hello[1] == 0 || hello[1] >= (uint32_t)world || c->fds[hello[1]] >= 0) {
        close(fd);                      // not one of ours (stale client, port scanner): ignore
        continue;
      }
      c->fds[hello[1]] = fd;
      ++have;
    }
    const uint32_t ok = kMagic;
    for (int r = 1; r < world; ++r)
      if (!send_all(c->fds[r], &ok, sizeof(ok), timeout_ms)) {
        bcg_comm_destroy(c);
        return bcg_set_error(BCG_ERR_COMM, "rank 0: handshake with rank %d failed", r);
      }
  } else {
    int fd = -1;
    for (;;) {                          // rank 0 may not be listening yet: retry until the deadline
      fd = socket(AF_INET, SOCK_STREAM, 0);
      if (fd >= 0 && connect(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) == 0) break;
      if (fd >= 0) close(fd);
      fd = -1;
      if (now_ms() > t_end) break;
      usleep(20000);
    }
    if (fd < 0) {
      delete c;
      return bcg_set_error(BCG_ERR_COMM, "rank %d cannot reach rank 0 at %s:%d within %d ms", rank, addr ? addr : "127.0.0.1",
                           port, timeout_ms);
    }
    tune(fd);
    const uint32_t hello[3] = {kMagic, (uint32_t)rank, (uint32_t)world};
    uint32_t ok = 0;
    if (!send_all(fd, hello, sizeof(hello), timeout_ms) || !recv_all(fd, &ok, sizeof(ok), timeout_ms) || ok != kMagic) {
      close(fd);
      delete c;
      return bcg_set_error(BCG_ERR_COMM, "rank %d: handshake with rank 0 failed", rank);
    }
    c->fds[0] = fd;
  }
  *out = c;
  return BCG_OK;
}

extern "C" int bcg_comm_destroy(bcg_comm* c) {
  if (!c) return BCG_OK;
  for (int fd : c->fds)
    if (fd >= 0) close(fd);
  if (c->listen_fd >= 0) close(c->listen_fd);
  delete c;
  return BCG_OK;
}

extern "C" int bcg_comm_rank(bcg_comm* c, int32_t* rank, int32_t* world) {
  if (!c) return bcg_set_error(BCG_ERR_ARG, "null comm");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return BCG_OK;
}

extern "C" int bcg_comm_allgather(bcg_comm* c, const void* send, int64_t bytes, void* recv) {
  if (!c || bytes < 0 || (bytes > 0 && (!send || !recv))) return bcg_set_error(BCG_ERR_ARG, "bad arguments");
  const size_t n = (size_t)bytes;
  const int W = c->world;
  if (W == 1) { if (n) memmove(recv, send, n); return BCG_OK; }
  char* out = static_cast<char*>(recv);
  if (c->rank == 0) {
    if (n) memmove(out, send, n);
    for (int r = 1; r < W; ++r)
      if (n && !recv_all(c->fds[r], out + (size_t)r * n, n, c->timeout_ms))
        return bcg_set_error(BCG_ERR_COMM, "all-gather: rank %d did not deliver within %d ms", r, c->timeout_ms);
    // zero-byte gathers still synchronise: one token each way
    uint32_t tok = kMagic;
    for (int r = 1; r < W; ++r) {
      if (!n && !recv_all(c->fds[r], &tok, sizeof(tok), c->timeout_ms)) return bcg_set_error(BCG_ERR_COMM, "barrier: rank %d is missing", r);
    }
    for (int r = 1; r < W; ++r) {
      const bool ok = n ? send_all(c->fds[r], out, n * W, c->timeout_ms) : send_all(c->fds[r], &tok, sizeof(tok), c->timeout_ms);
      if (!ok) return bcg_set_error(BCG_ERR_COMM, "all-gather: cannot reach rank %d", r);
    }
  } else {
    uint32_t tok = kMagic;
    const bool sent = n ? send_all(c->fds[0], send, n, c->timeout_ms) : send_all(c->fds[0], &tok, sizeof(tok), c->timeout_ms);
    const bool got = sent && (n ? recv_all(c->fds[0], out, n * W, c->timeout_ms) : recv_all(c->fds[0], &tok, sizeof(tok), c->timeout_ms));
    if (!got) return bcg_set_error(BCG_ERR_COMM, "all-gather: rank 0 did not answer within %d ms", c->timeout_ms);
  }
  return BCG_OK;
}

extern "C" int bcg_comm_barrier(bcg_comm* c) { return bcg_comm_allgather(c, nullptr, 0, nullptr); }

// op: 0 = sum, 1 = max.  Every rank reduces the gathered contributions in rank order: bit-identical results.
extern "C" int bcg_comm_allreduce_f64(bcg_comm* c, double* data, int64_t n, int32_t op) {
  if (!c || n < 0 || (n > 0 && !data) || (op != 0 && op != 1)) return bcg_set_error(BCG_ERR_ARG, "bad arguments");
  if (c->world == 1 || n == 0) return BCG_OK;
  c->buf.resize((size_t)n * sizeof(double) * c->world);
  const int rc = bcg_comm_allgather(c, data, n * (int64_t)sizeof(double), c->buf.data());
  if (rc != BCG_OK) return rc;
  const double* all = reinterpret_cast<const double*>(c->buf.data());
  for (int64_t i = 0; i < n; ++i) {
    double v = all[i];
    for (int r = 1; r < c->world; ++r) {
      const double x = all[(size_t)r * n + i];
      v = op == 0 ? v + x : fmax(v, x);
    }
    data[i] = v;
  }
  return BCG_OK;
}
