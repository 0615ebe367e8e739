// C ABI of libquda_b200.so (declared in include/b200_dslash.h).  Plain pointers and sizes only.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "launch.h"

#ifndef B2_TMA_DEFAULT
#define B2_TMA_DEFAULT 0
#endif

namespace b200
{
  static thread_local char g_err[512] = "";
  static std::atomic<long> g_launches {0};

  int set_error(int code, const char *fmt, ...)
  {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
  }

  int check_cuda(cudaError_t e, const char *what)
  {
    if (e == cudaSuccess) return B200_SUCCESS;
    return set_error(B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }

  void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

  static int require_device()
  {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      return set_error(B200_ERR_NO_DEVICE, "no CUDA device available: libquda_b200 has no CPU path (%s)",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return 0;
  }

  // B200_TMA=0|1|2 selects the TMA-staged marching kernel (2: fail instead of falling back when a shape is not served);
  // B200_TMA_TILE="ty tz", B200_TMA_GRID, B200_TMA_LINKS (n shared-memory stages, -1 register
  // stream), B200_TMA_L2PF (L2 prefetch look-ahead in items, -1 off), B200_TMA_PREFETCH, B200_TMA_RINGS="centre halo" tune it.  Read on every call (a handful of getenv lookups) so
  // that tests and tuning scripts can switch between calls.
  static void tma_knobs(LaunchRequest &rq)
  {
    const char *e = getenv("B200_TMA");
    rq.tma = e ? atoi(e) : B2_TMA_DEFAULT;
    if (!rq.tma) return;
    if ((e = getenv("B200_TMA_TILE"))) sscanf(e, "%d %d", &rq.tma_ty, &rq.tma_tz);
    if ((e = getenv("B200_TMA_GRID"))) rq.tma_grid = atoi(e);
    if ((e = getenv("B200_TMA_LINKS"))) rq.tma_link_slots = atoi(e);
    if ((e = getenv("B200_TMA_PREFETCH"))) rq.tma_prefetch = atoi(e);
    if ((e = getenv("B200_TMA_L2PF"))) rq.tma_l2_prefetch = atoi(e);
    if ((e = getenv("B200_TMA_RINGS"))) sscanf(e, "%d %d", &rq.tma_center_slots, &rq.tma_halo_slots);
  }

} // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error(void) { return g_err; }
int b200_abi_version(void) { return B200_ABI_VERSION; }
long b200_launch_count(void) { return g_launches.load(); }
void b200_reset_launch_count(void) { g_launches.store(0); }

size_t b200_ghost_face_bytes(int precision, const int X[4], int dim)
{
  return 2 * ghost_parity_bytes(precision, X, dim);
}

int b200_dslash_apply(const b200_dslash_args *a)
{
  if (int rc = require_device()) return rc;
  LaunchRequest rq;
  bool nothing_to_do = false;
  if (int rc = make_request(rq, a, nothing_to_do)) return rc;
  if (nothing_to_do) return B200_SUCCESS;
  tma_knobs(rq);
  if (rq.tma && rq.kernel == B200_KERNEL_AUTO && a->precision != B200_HALF) {
    // TMA-staged marching kernel (tma_kernel.cuh) for the shapes it serves; anything else falls through
    const int rc = a->precision == B200_DOUBLE ? launch_tma_precision<PrecF64>(rq) : launch_tma_precision<PrecF32>(rq);
    if (rc != kTmaSkip) return rc;
    if (rq.tma > 1) return set_error(B200_ERR_UNSUPPORTED, "B200_TMA=2: shape not served by the TMA kernel");
  }
  switch (a->precision) {
  case B200_DOUBLE: return launch_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_precision<PrecF32>(rq);
  case B200_HALF: return launch_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", a->precision);
}

int b200_dslash_apply_multi(const b200_dslash_args *a, int n_src, const b200_spinor *out, const b200_spinor *in,
                            const b200_spinor *x)
{
  if (int rc = require_device()) return rc;
  MrhsRequest rq;
  bool batched = false;
  if (int rc = make_mrhs_request(rq, a, n_src, out, in, x, batched)) return rc;
  if (!batched) { // other kernel selectors / operators: source by source, each on its own ghost slab
    for (int i = 0; i < n_src; i++) {
      const b200_dslash_args one = source_args(*a, i, out, in, x);
      if (int rc = b200_dslash_apply(&one)) return rc;
    }
    return B200_SUCCESS;
  }
  if (rq.interior_box && a->kernel == B200_KERNEL_AUTO) {
    // partitioned lattice, whole operator: the boundary tiles of every source (they wait for the batch's arrival counters
    // and read the source's own ghost slab) after the batched interior below -- the order the single-source AUTO path uses
    b200_dslash_args interior = *a;
    interior.kernel = B200_KERNEL_INTERIOR_TILES;
    if (int rc = b200_dslash_apply_multi(&interior, n_src, out, in, x)) return rc;
    for (int i = 0; i < n_src; i++) {
      b200_dslash_args one = source_args(*a, i, out, in, x);
      one.kernel = B200_KERNEL_BOUNDARY_TILES;
      if (int rc = b200_dslash_apply(&one)) return rc;
    }
    return B200_SUCCESS;
  }
  // tuning knobs (defaults are the measured best): B200_MRHS_MODE=thread|cta|auto, B200_MRHS_BATCH (sources per thread),
  // B200_MRHS_CTA_SOURCES (sources per CTA), B200_MRHS_L1=0|1 (link loads allocate in L1)
  if (const char *e = getenv("B200_MRHS_MODE")) rq.mode = (strcmp(e, "cta") == 0) ? 1 : (strcmp(e, "thread") == 0 ? 0 : -1);
  if (const char *e = getenv("B200_MRHS_BATCH")) rq.max_batch = atoi(e);
  if (const char *e = getenv("B200_MRHS_CTA_SOURCES")) rq.cta_sources = atoi(e);
  if (const char *e = getenv("B200_MRHS_L1")) rq.l1_links = atoi(e) ? 1 : 0;
  if (const char *e = getenv("B200_MRHS_CTA_CFG")) rq.cta_cfg = atoi(e);
  switch (a->precision) {
  case B200_DOUBLE: return launch_mrhs_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_mrhs_precision<PrecF32>(rq);
  case B200_HALF: return launch_mrhs_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", a->precision);
}

int b200_clover_apply(const b200_spinor *out, const b200_spinor *in, const b200_clover *A, int precision, int inverse,
                      int parity, void *stream)
{
  if (!out || !in || !A || !out->v || !in->v || !A->clover) return set_error(B200_ERR_INVALID, "null argument");
  if (int rc = require_device()) return rc;
  if (out->n_parity != 1 || in->n_parity != 1) return set_error(B200_ERR_INVALID, "ApplyClover acts on single-parity fields");
  if (parity != 0 && parity != 1) return set_error(B200_ERR_INVALID, "parity %d", parity);
  // inverse with a static-inverse build: the caller passes the A^-1 field and the kernel multiplies (same path as a
  // forward apply); with dynamic_inverse the field holds A and the kernel solves by Cholesky
  CloverRequest rq;
  rq.out = out->v;
  rq.out_norm = out->norm;
  rq.in = in->v;
  rq.in_norm = in->norm;
  rq.A = *A;
  rq.volume_cb = out->volume_cb;
  rq.inverse = inverse ? 1 : 0;
  rq.parity = parity;
  rq.stream = stream;
  switch (precision) {
  case B200_DOUBLE: return launch_clover_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_clover_precision<PrecF32>(rq);
  case B200_HALF: return launch_clover_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", precision);
}

// a, b of out = a (1 + i b gamma5) in for the singlet twist (dslash_gamma_helper.cuh:55-62)
static void twist_coefficients(double &a, double &b, double kappa, double mu, int dagger, int inverse)
{
  if (!inverse) {
    b = 2.0 * kappa * mu;
    a = 1.0;
  } else {
    b = -2.0 * kappa * mu;
    a = 1.0 / (1.0 + b * b);
  }
  if (dagger) b = -b;
}

int b200_twist_gamma5(const b200_spinor *out, const b200_spinor *in, int precision, double kappa, double mu, int dagger,
                      int inverse, void *stream)
{
  if (!out || !in || !out->v || !in->v) return set_error(B200_ERR_INVALID, "b200_twist_gamma5: null argument");
  if (int rc = require_device()) return rc;
  if (out->n_parity != 1 || in->n_parity != 1 || out->volume_cb != in->volume_cb)
    return set_error(B200_ERR_INVALID, "b200_twist_gamma5 acts on single-parity fields of equal volume");
  TwistRequest rq;
  rq.out = out->v;
  rq.out_norm = out->norm;
  rq.in = in->v;
  rq.in_norm = in->norm;
  rq.volume_cb = out->volume_cb;
  twist_coefficients(rq.a, rq.b, kappa, mu, dagger, inverse);
  rq.stream = stream;
  switch (precision) {
  case B200_DOUBLE: return launch_twist_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_twist_precision<PrecF32>(rq);
  case B200_HALF: return launch_twist_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", precision);
}

int b200_copy_spinor(const b200_spinor *native, int native_precision, void *host_order, int host_precision, int to_native,
                     void *stream)
{
  if (!native || !native->v || !host_order) return set_error(B200_ERR_INVALID, "b200_copy_spinor: null argument");
  if (int rc = require_device()) return rc;
  if (native->n_parity != 1) return set_error(B200_ERR_INVALID, "b200_copy_spinor converts one parity block per call");
  if (host_precision != B200_DOUBLE && host_precision != B200_SINGLE)
    return set_error(B200_ERR_INVALID, "host order precision must be 8 or 4 (got %d)", host_precision);
  CopyRequest rq;
  rq.native = native->v;
  rq.native_norm = native->norm;
  rq.host = host_order;
  rq.volume_cb = native->volume_cb;
  rq.host_precision = host_precision;
  rq.to_native = to_native ? 1 : 0;
  rq.stream = stream;
  switch (native_precision) {
  case B200_DOUBLE: return launch_copy_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_copy_precision<PrecF32>(rq);
  case B200_HALF: return launch_copy_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", native_precision);
}

int b200_copy_gauge(const b200_gauge *native, int native_precision, const int X[4], void *const qdp[4],
                    void *const ghost_links[4], int host_precision, void *stream)
{
  if (!native || !native->gauge || !X || !qdp) return set_error(B200_ERR_INVALID, "b200_copy_gauge: null argument");
  if (int rc = require_device()) return rc;
  if (host_precision != B200_DOUBLE && host_precision != B200_SINGLE)
    return set_error(B200_ERR_INVALID, "host gauge precision must be 8 or 4 (got %d)", host_precision);
  GaugeCopyRequest rq;
  rq.native = *native;
  for (int d = 0; d < 4; d++) {
    if (X[d] < 2 || (X[d] & 1)) return set_error(B200_ERR_INVALID, "X[%d]=%d must be even and >= 2", d, X[d]);
    rq.X[d] = X[d];
    if (!qdp[d]) return set_error(B200_ERR_INVALID, "qdp[%d] is NULL", d);
    rq.qdp[d] = qdp[d];
    rq.ghost[d] = ghost_links ? ghost_links[d] : nullptr;
  }
  if (native_precision == B200_HALF && native->reconstruct == 18 && !(native->link_max > 0.0))
    return set_error(B200_ERR_INVALID, "half recon-18 needs link_max > 0");
  rq.host_precision = host_precision;
  rq.stream = stream;
  switch (native_precision) {
  case B200_DOUBLE: return launch_gauge_copy_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_gauge_copy_precision<PrecF32>(rq);
  case B200_HALF: return launch_gauge_copy_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", native_precision);
}

int b200_copy_clover(const b200_clover *native, int native_precision, const int X[4], const void *packed, int host_precision,
                     void *stream)
{
  if (!native || !native->clover || !X || !packed) return set_error(B200_ERR_INVALID, "b200_copy_clover: null argument");
  if (int rc = require_device()) return rc;
  if (host_precision != B200_DOUBLE && host_precision != B200_SINGLE)
    return set_error(B200_ERR_INVALID, "host clover precision must be 8 or 4 (got %d)", host_precision);
  if (native_precision == B200_HALF && !(native->max_element > 0.0))
    return set_error(B200_ERR_INVALID, "half precision clover needs max_element > 0");
  CloverCopyRequest rq;
  rq.native = *native;
  for (int d = 0; d < 4; d++) rq.X[d] = X[d];
  rq.packed = packed;
  rq.host_precision = host_precision;
  rq.stream = stream;
  switch (native_precision) {
  case B200_DOUBLE: return launch_clover_copy_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_clover_copy_precision<PrecF32>(rq);
  case B200_HALF: return launch_clover_copy_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", native_precision);
}

int b200_comm_alloc(void **ptr, size_t bytes)
{
  if (!ptr || bytes == 0) return set_error(B200_ERR_INVALID, "b200_comm_alloc: null pointer or zero size");
  if (int rc = require_device()) return rc;
  if (int rc = check_cuda(cudaMalloc(ptr, bytes), "cudaMalloc")) return rc;
  return check_cuda(cudaMemset(*ptr, 0, bytes), "cudaMemset");
}

int b200_comm_free(void *ptr) { return check_cuda(cudaFree(ptr), "cudaFree"); }

int b200_ipc_get_handle(void *ptr, unsigned char handle[B200_IPC_HANDLE_BYTES])
{
  static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (int rc = check_cuda(cudaIpcGetMemHandle(&h, ptr), "cudaIpcGetMemHandle")) return rc;
  memcpy(handle, &h, sizeof(h));
  return B200_SUCCESS;
}

int b200_ipc_open_handle(const unsigned char handle[B200_IPC_HANDLE_BYTES], void **peer_ptr)
{
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return check_cuda(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int b200_ipc_close_handle(void *peer_ptr) { return check_cuda(cudaIpcCloseMemHandle(peer_ptr), "cudaIpcCloseMemHandle"); }

int b200_comm_copy(void *dst, const void *src, size_t bytes)
{
  return check_cuda(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault), "cudaMemcpy");
}

static int make_pack_request(PackRequest &rq, const b200_pack_args *a)
{
  if (!a || !a->in.v) return set_error(B200_ERR_INVALID, "null argument");
  if (a->abi_version != B200_ABI_VERSION) return set_error(B200_ERR_INVALID, "ABI version mismatch");
  memcpy(rq.X, a->X, sizeof(rq.X));
  rq.parity = a->parity;
  rq.dagger = a->dagger ? 1 : 0;
  rq.in = a->in.v;
  rq.in_norm = a->in.norm;
  for (int d = 0; d < 4; d++) {
    rq.comm_dim[d] = a->comm_dim[d];
    for (int dir = 0; dir < 2; dir++) {
      rq.dst[d][dir] = a->dst[d][dir];
      rq.dst_norm[d][dir] = a->dst_norm[d][dir];
      rq.signal[d][dir] = a->signal[d][dir];
      if (a->comm_dim[d] && !a->dst[d][dir]) return set_error(B200_ERR_INVALID, "dst[%d][%d] is NULL", d, dir);
    }
  }
  rq.block_counter = a->block_counter;
  rq.seq = a->seq;
  rq.stream = a->stream;
  return 0;
}

int b200_pack_ghost(const b200_pack_args *a)
{
  PackRequest rq;
  if (int rc = make_pack_request(rq, a)) return rc;
  if (int rc = require_device()) return rc;
  switch (a->precision) {
  case B200_DOUBLE: return launch_pack_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_pack_precision<PrecF32>(rq);
  case B200_HALF: return launch_pack_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", a->precision);
}

int b200_pack_ghost_multi(const b200_pack_args *a, int n_src, const b200_spinor *in, const size_t dst_stride[4])
{
  if (n_src < 1 || n_src > B200_MAX_MULTI_RHS) return set_error(B200_ERR_INVALID, "n_src %d not in [1, %d]", n_src, B200_MAX_MULTI_RHS);
  if (!a || !in || !dst_stride) return set_error(B200_ERR_INVALID, "null argument");
  b200_pack_args first = *a;
  first.in = in[0];
  PackRequest rq;
  if (int rc = make_pack_request(rq, &first)) return rc;
  PackBatchRequest batch {};
  batch.n_src = n_src;
  for (int s = 0; s < n_src; s++) {
    if (!in[s].v) return set_error(B200_ERR_INVALID, "source %d is null", s);
    if (in[s].n_parity != 1) return set_error(B200_ERR_INVALID, "the batched pack takes single-parity sources");
    batch.in[s] = in[s].v;
    batch.in_norm[s] = in[s].norm;
  }
  for (int d = 0; d < 4; d++) batch.dst_stride[d] = dst_stride[d];
  if (int rc = check_src_stride("dst_stride", n_src, a->precision, a->X, a->comm_dim, dst_stride, 1)) return rc;
  if (int rc = require_device()) return rc;
  switch (a->precision) {
  case B200_DOUBLE: return launch_pack_multi_precision<PrecF64>(rq, batch);
  case B200_SINGLE: return launch_pack_multi_precision<PrecF32>(rq, batch);
  case B200_HALF: return launch_pack_multi_precision<PrecH16>(rq, batch);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", a->precision);
}

int b200_dslash_apply_fused(const b200_dslash_args *a, const b200_pack_args *p)
{
  if (int rc = require_device()) return rc;
  LaunchRequest rq;
  bool nothing_to_do = false;
  if (int rc = make_request(rq, a, nothing_to_do)) return rc;
  PackRequest pk;
  if (int rc = make_pack_request(pk, p)) return rc;
  if (a->kernel != B200_KERNEL_AUTO && a->kernel != B200_KERNEL_INTERIOR_TILES && a->kernel != B200_KERNEL_BOUNDARY_TILES)
    return set_error(B200_ERR_INVALID, "fused Dslash: kernel must be AUTO (all roles), INTERIOR_TILES (interior role) or BOUNDARY_TILES (pack + boundary roles)");
  if (p->in.v != a->in.v || p->precision != a->precision || p->parity != 1 - a->parity || (p->dagger != 0) != (a->dagger != 0))
    return set_error(B200_ERR_INVALID, "pack arguments do not describe the faces of this Dslash's input");
  if (a->out.n_parity != 1) return set_error(B200_ERR_INVALID, "the fused Dslash works on single-parity fields");
  bool any = false;
  for (int d = 0; d < 4; d++) {
    if ((p->comm_dim[d] != 0) != (a->halo.comm_dim[d] != 0)) return set_error(B200_ERR_INVALID, "pack / halo partitioning differ in dimension %d", d);
    if (p->X[d] != a->X[d]) return set_error(B200_ERR_INVALID, "pack / Dslash lattice extents differ");
    any |= a->halo.comm_dim[d] != 0;
  }
  if (!any) return a->kernel == B200_KERNEL_BOUNDARY_TILES ? B200_SUCCESS : b200_dslash_apply(a); // nothing partitioned: the plain launch
  if (a->halo.seq != p->seq) return set_error(B200_ERR_INVALID, "pack and halo carry different sequence numbers");
  rq.fused_pack = &pk;
  switch (a->precision) {
  case B200_DOUBLE: return launch_precision<PrecF64>(rq);
  case B200_SINGLE: return launch_precision<PrecF32>(rq);
  case B200_HALF: return launch_precision<PrecH16>(rq);
  }
  return set_error(B200_ERR_INVALID, "precision %d not in {8,4,2}", a->precision);
}
}
