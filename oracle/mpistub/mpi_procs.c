/*
 * N-rank implementation of oracle/mpistub/mpi.h for the standalone reference driver
 * (oracle/ref_driver.cpp) -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * MPI_Init reads TACSB200_MPI_NP, creates a full mesh of AF_UNIX socket pairs and forks NP-1
 * children; every process then runs the unmodified reference as one MPI rank. Point-to-point
 * messages are framed (tag, bytes) on the pair's stream; sends are buffered (copied and queued) and a
 * small progress engine moves queued sends / incoming messages with non-blocking I/O, so the
 * Isend/Irecv/Waitall patterns of TACSBVecDistribute and TACSMatDistribute cannot deadlock.
 * Collectives are linear algorithms over the same point-to-point layer with reserved tags.
 * With TACSB200_MPI_NP unset or 1 this behaves like mpi_single.c.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define MAXP 64
#define TAG_COLL (-1000)

struct tacsb200_mpi_file {
  FILE *fp;
  MPI_Offset disp;
};

typedef struct msg {
  int tag, nbytes;
  char *data;
  struct msg *next;
} msg_t;

typedef struct sendq {
  char *buf; /* header + payload */
  size_t len, off;
  struct sendq *next;
} sendq_t;

typedef struct {
  int active, is_recv, done;
  int peer, tag;
  void *buf;
  int nbytes;
  MPI_Status st;
} req_t;

static int g_rank = 0, g_size = 1, g_init = 0;
static int g_fd[MAXP];
static msg_t *g_inbox[MAXP], *g_inbox_tail[MAXP];
static sendq_t *g_out[MAXP], *g_out_tail[MAXP];
/* partial incoming message per peer */
static struct {
  int have_hdr;
  int hdr[2];
  size_t got;
  char *data;
} g_in[MAXP];
static req_t g_req[1 << 16];
static int g_nreq = 0;
static pid_t g_children[MAXP];

static size_t esize(MPI_Datatype t) { return (size_t)(t & 0xff); }

static void die(const char *what) {
  fprintf(stderr, "[oracle mpi_procs rank %d] %s: %s\n", g_rank, what, strerror(errno));
  abort();
}

static void set_nonblock(int fd) {
  int fl = fcntl(fd, F_GETFL, 0);
  fcntl(fd, F_SETFL, fl | O_NONBLOCK);
}

/* one pass of non-blocking I/O; if block != 0 wait in poll() until something can move */
static void progress(int block) {
  struct pollfd pf[MAXP];
  int map[MAXP], n = 0;
  for (int p = 0; p < g_size; p++) {
    if (p == g_rank) continue;
    pf[n].fd = g_fd[p];
    pf[n].events = POLLIN | (g_out[p] ? POLLOUT : 0);
    pf[n].revents = 0;
    map[n++] = p;
  }
  if (n == 0) return;
  if (poll(pf, n, block ? 1000 : 0) < 0 && errno != EINTR) die("poll");
  for (int k = 0; k < n; k++) {
    int p = map[k];
    if (pf[k].revents & POLLOUT) {
      while (g_out[p]) {
        sendq_t *s = g_out[p];
        ssize_t w = write(g_fd[p], s->buf + s->off, s->len - s->off);
        if (w < 0) {
          if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
          die("write");
        }
        s->off += (size_t)w;
        if (s->off < s->len) break;
        g_out[p] = s->next;
        if (!g_out[p]) g_out_tail[p] = NULL;
        free(s->buf);
        free(s);
      }
    }
    if (pf[k].revents & (POLLIN | POLLHUP)) {
      for (;;) {
        if (!g_in[p].have_hdr) {
          ssize_t r = read(g_fd[p], (char *)g_in[p].hdr + g_in[p].got, sizeof(g_in[p].hdr) - g_in[p].got);
          if (r < 0) {
            if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
            die("read");
          }
          if (r == 0) break; /* peer closed */
          g_in[p].got += (size_t)r;
          if (g_in[p].got < sizeof(g_in[p].hdr)) break;
          g_in[p].have_hdr = 1;
          g_in[p].got = 0;
          g_in[p].data = (char *)malloc(g_in[p].hdr[1] > 0 ? (size_t)g_in[p].hdr[1] : 1);
        }
        size_t want = (size_t)g_in[p].hdr[1];
        if (g_in[p].got < want) {
          ssize_t r = read(g_fd[p], g_in[p].data + g_in[p].got, want - g_in[p].got);
          if (r < 0) {
            if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
            die("read");
          }
          if (r == 0) break;
          g_in[p].got += (size_t)r;
          if (g_in[p].got < want) break;
        }
        msg_t *m = (msg_t *)malloc(sizeof(msg_t));
        m->tag = g_in[p].hdr[0];
        m->nbytes = g_in[p].hdr[1];
        m->data = g_in[p].data;
        m->next = NULL;
        if (g_inbox_tail[p]) g_inbox_tail[p]->next = m;
        else g_inbox[p] = m;
        g_inbox_tail[p] = m;
        g_in[p].have_hdr = 0;
        g_in[p].got = 0;
        g_in[p].data = NULL;
      }
    }
  }
}

static void post_send(const void *buf, int nbytes, int dest, int tag) {
  if (dest == g_rank) {
    msg_t *m = (msg_t *)malloc(sizeof(msg_t));
    m->tag = tag;
    m->nbytes = nbytes;
    m->data = (char *)malloc(nbytes > 0 ? (size_t)nbytes : 1);
    memcpy(m->data, buf, (size_t)nbytes);
    m->next = NULL;
    if (g_inbox_tail[dest]) g_inbox_tail[dest]->next = m;
    else g_inbox[dest] = m;
    g_inbox_tail[dest] = m;
    return;
  }
  sendq_t *s = (sendq_t *)malloc(sizeof(sendq_t));
  s->len = 2 * sizeof(int) + (size_t)nbytes;
  s->buf = (char *)malloc(s->len);
  ((int *)s->buf)[0] = tag;
  ((int *)s->buf)[1] = nbytes;
  memcpy(s->buf + 2 * sizeof(int), buf, (size_t)nbytes);
  s->off = 0;
  s->next = NULL;
  if (g_out_tail[dest]) g_out_tail[dest]->next = s;
  else g_out[dest] = s;
  g_out_tail[dest] = s;
  progress(0);
}

/* first queued message from `src` whose tag matches; removes it from the queue */
static msg_t *take(int src, int tag) {
  msg_t *prev = NULL;
  for (msg_t *m = g_inbox[src]; m; prev = m, m = m->next) {
    if (tag == MPI_ANY_TAG ? (m->tag > TAG_COLL) : (m->tag == tag)) {
      if (prev) prev->next = m->next;
      else g_inbox[src] = m->next;
      if (g_inbox_tail[src] == m) g_inbox_tail[src] = prev;
      return m;
    }
  }
  return NULL;
}

static int try_recv(void *buf, int nbytes, int src, int tag, MPI_Status *st) {
  int lo = src == MPI_ANY_SOURCE ? 0 : src, hi = src == MPI_ANY_SOURCE ? g_size : src + 1;
  for (int p = lo; p < hi; p++) {
    msg_t *m = take(p, tag);
    if (!m) continue;
    if (m->nbytes > nbytes) {
      fprintf(stderr, "[oracle mpi_procs rank %d] message of %d bytes from %d exceeds buffer %d\n", g_rank,
              m->nbytes, p, nbytes);
      abort();
    }
    memcpy(buf, m->data, (size_t)m->nbytes);
    if (st) {
      st->MPI_SOURCE = p;
      st->MPI_TAG = m->tag;
      st->MPI_ERROR = 0;
      st->nbytes_ = m->nbytes;
    }
    free(m->data);
    free(m);
    return 1;
  }
  return 0;
}

static void blocking_recv(void *buf, int nbytes, int src, int tag, MPI_Status *st) {
  while (!try_recv(buf, nbytes, src, tag, st)) progress(1);
}

static void flush_sends(void) {
  for (;;) {
    int pending = 0;
    for (int p = 0; p < g_size; p++)
      if (g_out[p]) pending = 1;
    if (!pending) return;
    progress(1);
  }
}

/* ---- init / finalize -------------------------------------------------------------------- */
int MPI_Init(int *argc, char ***argv) {
  (void)argc; (void)argv;
  if (g_init) return MPI_SUCCESS;
  g_init = 1;
  const char *np = getenv("TACSB200_MPI_NP");
  g_size = np ? atoi(np) : 1;
  if (g_size < 1) g_size = 1;
  if (g_size > MAXP) g_size = MAXP;
  if (g_size == 1) return MPI_SUCCESS;
  static int pairs[MAXP][MAXP][2];
  for (int i = 0; i < g_size; i++)
    for (int j = i + 1; j < g_size; j++)
      if (socketpair(AF_UNIX, SOCK_STREAM, 0, pairs[i][j]) < 0) die("socketpair");
  fflush(stdout);
  fflush(stderr);
  g_rank = 0;
  for (int r = 1; r < g_size; r++) {
    pid_t pid = fork();
    if (pid < 0) die("fork");
    if (pid == 0) {
      g_rank = r;
      break;
    }
    g_children[r] = pid;
  }
  for (int i = 0; i < g_size; i++)
    for (int j = i + 1; j < g_size; j++) {
      if (i == g_rank) { g_fd[j] = pairs[i][j][0]; close(pairs[i][j][1]); }
      else if (j == g_rank) { g_fd[i] = pairs[i][j][1]; close(pairs[i][j][0]); }
      else { close(pairs[i][j][0]); close(pairs[i][j][1]); }
    }
  for (int p = 0; p < g_size; p++)
    if (p != g_rank) {
      set_nonblock(g_fd[p]);
      int sz = 4 << 20;
      setsockopt(g_fd[p], SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
      setsockopt(g_fd[p], SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
    }
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (g_size > 1) {
    MPI_Barrier(MPI_COMM_WORLD);
    flush_sends();
    if (g_rank == 0) {
      for (int r = 1; r < g_size; r++) {
        int status = 0;
        waitpid(g_children[r], &status, 0);
      }
    } else {
      fflush(stdout);
      fflush(stderr);
      _exit(0);
    }
  }
  return MPI_SUCCESS;
}
int MPI_Initialized(int *flag) { *flag = g_init; return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) {
  (void)comm;
  fprintf(stderr, "[oracle mpi_procs rank %d] MPI_Abort(%d)\n", g_rank, code);
  abort();
  return code;
}
static int comm_size(MPI_Comm c) { return c == MPI_COMM_SELF ? 1 : g_size; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm == MPI_COMM_SELF ? 0 : g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm_size(comm); return MPI_SUCCESS; }
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *result) {
  *result = (a == b) ? MPI_IDENT : (comm_size(a) == comm_size(b) ? MPI_CONGRUENT : MPI_UNEQUAL);
  return MPI_SUCCESS;
}
double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- point to point --------------------------------------------------------------------- */
int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
  (void)c;
  post_send(b, (int)(n * esize(t)), d, tag);
  return MPI_SUCCESS;
}
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) {
  (void)c;
  blocking_recv(b, (int)(n * esize(t)), s, tag, st);
  return MPI_SUCCESS;
}
static int new_req(void) {
  for (int k = 0; k < g_nreq; k++)
    if (!g_req[k].active) return k;
  if (g_nreq >= (int)(sizeof(g_req) / sizeof(g_req[0]))) die("too many requests");
  return g_nreq++;
}
int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) {
  (void)c;
  post_send(b, (int)(n * esize(t)), d, tag); /* buffered: complete on return */
  int k = new_req();
  g_req[k].active = 1;
  g_req[k].is_recv = 0;
  g_req[k].done = 1;
  *r = k + 1;
  return MPI_SUCCESS;
}
int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r) {
  (void)c;
  int k = new_req();
  g_req[k].active = 1;
  g_req[k].is_recv = 1;
  g_req[k].done = 0;
  g_req[k].peer = s;
  g_req[k].tag = tag;
  g_req[k].buf = b;
  g_req[k].nbytes = (int)(n * esize(t));
  *r = k + 1;
  return MPI_SUCCESS;
}
static int test_req(int k) {
  req_t *q = &g_req[k];
  if (!q->done && q->is_recv) q->done = try_recv(q->buf, q->nbytes, q->peer, q->tag, &q->st);
  return q->done;
}
int MPI_Wait(MPI_Request *r, MPI_Status *s) {
  if (*r <= 0) return MPI_SUCCESS;
  int k = *r - 1;
  while (!test_req(k)) progress(1);
  if (s && g_req[k].is_recv) *s = g_req[k].st;
  g_req[k].active = 0;
  *r = 0;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) {
  for (int i = 0; i < n; i++) MPI_Wait(&r[i], s ? &s[i] : NULL);
  return MPI_SUCCESS;
}
int MPI_Waitany(int n, MPI_Request *r, int *index, MPI_Status *s) {
  int live = 0;
  for (int i = 0; i < n; i++)
    if (r[i] > 0) live = 1;
  if (!live) {
    *index = MPI_UNDEFINED;
    return MPI_SUCCESS;
  }
  for (;;) {
    for (int i = 0; i < n; i++) {
      if (r[i] <= 0) continue;
      if (test_req(r[i] - 1)) {
        if (s && g_req[r[i] - 1].is_recv) *s = g_req[r[i] - 1].st;
        g_req[r[i] - 1].active = 0;
        r[i] = 0;
        *index = i;
        return MPI_SUCCESS;
      }
    }
    progress(1);
  }
}
int MPI_Probe(int src, int tag, MPI_Comm c, MPI_Status *st) {
  (void)c;
  for (;;) {
    int lo = src == MPI_ANY_SOURCE ? 0 : src, hi = src == MPI_ANY_SOURCE ? g_size : src + 1;
    for (int p = lo; p < hi; p++)
      for (msg_t *m = g_inbox[p]; m; m = m->next)
        if (tag == MPI_ANY_TAG ? (m->tag > TAG_COLL) : (m->tag == tag)) {
          if (st) { st->MPI_SOURCE = p; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->nbytes_ = m->nbytes; }
          return MPI_SUCCESS;
        }
    progress(1);
  }
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) {
  *count = st ? (int)(st->nbytes_ / (int)esize(t)) : 0;
  return MPI_SUCCESS;
}

/* ---- collectives (linear, through rank `root`) ----------------------------------------------- */
static int g_coll_seq = 0;
static int coll_tag(void) { return TAG_COLL - 1 - (g_coll_seq++ % 1000000); }

int MPI_Barrier(MPI_Comm c) {
  if (comm_size(c) == 1) return MPI_SUCCESS;
  int tag = coll_tag(), token = 0;
  if (g_rank == 0) {
    for (int p = 1; p < g_size; p++) blocking_recv(&token, sizeof(int), p, tag, NULL);
    for (int p = 1; p < g_size; p++) post_send(&token, sizeof(int), p, tag);
  } else {
    post_send(&token, sizeof(int), 0, tag);
    blocking_recv(&token, sizeof(int), 0, tag, NULL);
  }
  return MPI_SUCCESS;
}
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  if (comm_size(c) == 1) return MPI_SUCCESS;
  int tag = coll_tag(), nb = (int)(n * esize(t));
  if (g_rank == root) {
    for (int p = 0; p < g_size; p++)
      if (p != root) post_send(b, nb, p, tag);
  } else {
    blocking_recv(b, nb, root, tag, NULL);
  }
  return MPI_SUCCESS;
}
static void combine(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op) {
  if (t == MPI_INT) {
    int *a = (int *)acc;
    const int *b = (const int *)in;
    for (int i = 0; i < n; i++) a[i] = op == MPI_SUM ? a[i] + b[i] : (op == MPI_MAX ? (a[i] > b[i] ? a[i] : b[i]) : (a[i] < b[i] ? a[i] : b[i]));
  } else if (t == MPI_DOUBLE) {
    double *a = (double *)acc;
    const double *b = (const double *)in;
    for (int i = 0; i < n; i++) a[i] = op == MPI_SUM ? a[i] + b[i] : (op == MPI_MAX ? (a[i] > b[i] ? a[i] : b[i]) : (a[i] < b[i] ? a[i] : b[i]));
  } else {
    fprintf(stderr, "[oracle mpi_procs] reduction on unsupported datatype %d\n", t);
    abort();
  }
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  size_t nb = n * esize(t);
  if (comm_size(c) == 1) {
    if (s != MPI_IN_PLACE && s != r) memmove(r, s, nb);
    return MPI_SUCCESS;
  }
  int tag = coll_tag();
  if (g_rank == root) {
    char *tmp = (char *)malloc(nb ? nb : 1);
    if (s != MPI_IN_PLACE && s != r) memmove(r, s, nb);
    for (int p = 0; p < g_size; p++) { /* rank order: deterministic sums */
      if (p == root) continue;
      blocking_recv(tmp, (int)nb, p, tag, NULL);
      combine(r, tmp, n, t, op);
    }
    free(tmp);
  } else {
    post_send(s == MPI_IN_PLACE ? r : s, (int)nb, root, tag);
  }
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  if (g_rank != 0 && s != MPI_IN_PLACE && s != r) memmove(r, s, n * esize(t));
  MPI_Reduce(g_rank == 0 ? s : MPI_IN_PLACE, r, n, t, op, 0, c);
  return MPI_Bcast(r, n, t, 0, c);
}
int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *displs, MPI_Datatype rt,
                int root, MPI_Comm c) {
  if (comm_size(c) == 1) {
    if (s != MPI_IN_PLACE) memmove((char *)r + displs[0] * esize(rt), s, sn * esize(st));
    return MPI_SUCCESS;
  }
  int tag = coll_tag();
  if (g_rank == root) {
    for (int p = 0; p < g_size; p++) {
      char *dst = (char *)r + displs[p] * esize(rt);
      if (p == root) {
        if (s != MPI_IN_PLACE) memmove(dst, s, sn * esize(st));
      } else {
        blocking_recv(dst, (int)(rc[p] * esize(rt)), p, tag, NULL);
      }
    }
  } else {
    post_send(s, (int)(sn * esize(st)), root, tag);
  }
  return MPI_SUCCESS;
}
int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  int rc[MAXP], d[MAXP];
  for (int p = 0; p < g_size; p++) { rc[p] = rn; d[p] = p * rn; }
  return MPI_Gatherv(s, sn, st, r, rc, d, rt, root, c);
}
int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  MPI_Gather(s, sn, st, r, rn, rt, 0, c);
  return MPI_Bcast(r, rn * comm_size(c), rt, 0, c);
}
int MPI_Scatterv(const void *s, const int *sc, const int *displs, MPI_Datatype st, void *r, int rn, MPI_Datatype rt,
                 int root, MPI_Comm c) {
  if (comm_size(c) == 1) {
    memmove(r, (const char *)s + displs[0] * esize(st), rn * esize(rt));
    return MPI_SUCCESS;
  }
  int tag = coll_tag();
  if (g_rank == root) {
    for (int p = 0; p < g_size; p++) {
      const char *src = (const char *)s + displs[p] * esize(st);
      if (p == root) memmove(r, src, rn * esize(rt));
      else post_send(src, (int)(sc[p] * esize(st)), p, tag);
    }
  } else {
    blocking_recv(r, (int)(rn * esize(rt)), root, tag, NULL);
  }
  return MPI_SUCCESS;
}
int MPI_Scatter(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  int sc[MAXP], d[MAXP];
  for (int p = 0; p < g_size; p++) { sc[p] = sn; d[p] = p * sn; }
  return MPI_Scatterv(s, sc, d, st, r, rn, rt, root, c);
}
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc,
                  const int *rd, MPI_Datatype rt, MPI_Comm c) {
  const int n = comm_size(c);
  if (n == 1) {
    memmove((char *)r + rd[0] * esize(rt), (const char *)s + sd[0] * esize(st), sc[0] * esize(st));
    return MPI_SUCCESS;
  }
  int tag = coll_tag();
  for (int p = 0; p < n; p++) post_send((const char *)s + sd[p] * esize(st), (int)(sc[p] * esize(st)), p, tag);
  for (int p = 0; p < n; p++) blocking_recv((char *)r + rd[p] * esize(rt), (int)(rc[p] * esize(rt)), p, tag, NULL);
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  int sc[MAXP], sd[MAXP], rc[MAXP], rd[MAXP];
  for (int p = 0; p < comm_size(c); p++) { sc[p] = sn; sd[p] = p * sn; rc[p] = rn; rd[p] = p * rn; }
  return MPI_Alltoallv(s, sc, sd, st, r, rc, rd, rt, c);
}

int MPI_Error_string(int code, char *str, int *len) {
  *len = snprintf(str, MPI_MAX_ERROR_STRING, "oracle mpi error %d", code);
  return MPI_SUCCESS;
}
int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op) { (void)fn; (void)commute; *op = 100; return MPI_SUCCESS; }
int MPI_Op_free(MPI_Op *op) { *op = 0; return MPI_SUCCESS; }

/* MPI-IO is outside the hot path; rank 0 writes a plain file so the symbols resolve */
int MPI_File_open(MPI_Comm c, const char *name, int mode, MPI_Info info, MPI_File *fp) {
  (void)c; (void)info;
  FILE *f = fopen(name, (mode & MPI_MODE_WRONLY) ? "wb" : "rb");
  if (!f) { *fp = NULL; return 1; }
  *fp = (MPI_File)calloc(1, sizeof(**fp));
  (*fp)->fp = f;
  return MPI_SUCCESS;
}
int MPI_File_close(MPI_File *fp) {
  if (*fp) { fclose((*fp)->fp); free(*fp); *fp = NULL; }
  return MPI_SUCCESS;
}
int MPI_File_set_view(MPI_File fp, MPI_Offset disp, MPI_Datatype et, MPI_Datatype ft, const char *rep, MPI_Info info) {
  (void)et; (void)ft; (void)rep; (void)info; fp->disp = disp; return MPI_SUCCESS;
}
int MPI_File_set_size(MPI_File fp, MPI_Offset size) { (void)fp; (void)size; return MPI_SUCCESS; }
int MPI_File_write(MPI_File fp, const void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; return fwrite(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
int MPI_File_read(MPI_File fp, void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; return fread(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
int MPI_File_write_at_all(MPI_File fp, MPI_Offset off, const void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; fseek(fp->fp, (long)(fp->disp + off * (MPI_Offset)esize(t)), SEEK_SET);
  return fwrite(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
int MPI_File_read_at_all(MPI_File fp, MPI_Offset off, void *b, int n, MPI_Datatype t, MPI_Status *st) {
  (void)st; fseek(fp->fp, (long)(fp->disp + off * (MPI_Offset)esize(t)), SEEK_SET);
  return fread(b, esize(t), n, fp->fp) == (size_t)n ? MPI_SUCCESS : 1;
}
