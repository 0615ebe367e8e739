// Implementation of the host-side operator / blas / solver layer (design notes and reference citations: dirac.h).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <cuda_runtime.h>

#include "dirac.h"

namespace b200
{
  namespace host
  {

    static void cuda_ok(cudaError_t e, const char *what)
    {
      if (e != cudaSuccess) throw Error(std::string(what) + ": " + cudaGetErrorString(e));
    }
    static void abi_ok(int rc)
    {
      if (rc != B200_SUCCESS) throw Error(b200_last_error());
    }
    static cudaStream_t cs(void *s) { return static_cast<cudaStream_t>(s); }

    // ------------------------------------------------------------------ fields
    static size_t parity_bytes_of(const int *X, int precision)
    {
      const size_t vcb = (size_t)X[0] * X[1] * X[2] * X[3] / 2;
      return vcb * 24 * precision + (precision == B200_HALF ? vcb * 4 : 0);
    }

    ColorSpinorField ColorSpinorField::wrap(void *v, const int *X, int precision, int n_parity)
    {
      ColorSpinorField f;
      f.v = v;
      for (int d = 0; d < 4; d++) f.X[d] = X[d];
      f.precision = precision;
      f.n_parity = n_parity;
      f.parity_bytes = parity_bytes_of(X, precision);
      return f;
    }

    ColorSpinorField ColorSpinorField::create(const int *X, int precision, int n_parity)
    {
      void *p = nullptr;
      const size_t bytes = parity_bytes_of(X, precision) * n_parity;
      cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc(ColorSpinorField)");
      cuda_ok(cudaMemset(p, 0, bytes), "cudaMemset(ColorSpinorField)");
      ColorSpinorField f = wrap(p, X, precision, n_parity);
      f.owned = std::shared_ptr<void>(p, [](void *q) { cudaFree(q); });
      return f;
    }

    ColorSpinorField ColorSpinorField::parity_view(int p) const
    {
      if (n_parity != 2) throw Error("parity_view needs a full (two-parity) field");
      ColorSpinorField f = wrap(static_cast<char *>(v) + p * parity_bytes, X, precision, 1);
      f.owned = owned;
      return f;
    }

    b200_spinor ColorSpinorField::desc() const
    {
      b200_spinor s;
      s.v = v;
      s.norm = nullptr;
      s.parity_stride_bytes = n_parity == 2 ? parity_bytes : 0;
      s.volume_cb = VolumeCB();
      s.n_parity = n_parity;
      return s;
    }

    // Scratch fields: an operator application needs one or two temporaries; cudaMalloc per call would serialise the
    // device, so released temporaries are parked and reused -- per stream, because a parked buffer may still be read by
    // kernels queued on the stream that used it last.
    namespace
    {
      struct ScratchPool {
        std::multimap<void *, ColorSpinorField> parked; // keyed by stream
        ColorSpinorField get(void *stream, const int *X, int precision, int n_parity)
        {
          auto range = parked.equal_range(stream);
          for (auto it = range.first; it != range.second; ++it) {
            const ColorSpinorField &f = it->second;
            if (f.precision == precision && f.n_parity == n_parity && f.X[0] == X[0] && f.X[1] == X[1] && f.X[2] == X[2]
                && f.X[3] == X[3]) {
              ColorSpinorField r = f;
              parked.erase(it);
              return r;
            }
          }
          return ColorSpinorField::create(X, precision, n_parity);
        }
        void put(void *stream, const ColorSpinorField &f) { parked.emplace(stream, f); }
      };
      ScratchPool &pool()
      {
        static ScratchPool p;
        return p;
      }
      struct Scratch {
        void *stream;
        ColorSpinorField f;
        Scratch(void *stream_, const ColorSpinorField &like, int n_parity) :
          stream(stream_), f(pool().get(stream_, like.X, like.precision, n_parity))
        {
        }
        Scratch(void *stream_, const int *X, int precision, int n_parity) : stream(stream_), f(pool().get(stream_, X, precision, n_parity)) { }
        ~Scratch() { pool().put(stream, f); }
        operator ColorSpinorField &() { return f; }
      };

      // fork / join events of the two-stream halo schedule, one pair per halo context
      struct StreamEvents {
        cudaEvent_t fork = nullptr, join = nullptr;
      };
      StreamEvents &events_of(CommContext *c)
      {
        static std::map<CommContext *, StreamEvents> m;
        StreamEvents &e = m[c];
        if (!e.fork) {
          cuda_ok(cudaEventCreateWithFlags(&e.fork, cudaEventDisableTiming), "event");
          cuda_ok(cudaEventCreateWithFlags(&e.join, cudaEventDisableTiming), "event");
        }
        return e;
      }
    } // namespace

    // ------------------------------------------------------------------ Apply*
    static void halo_fill(b200_halo &h, const int *comm_override, CommContext *comm)
    {
      memset(&h, 0, sizeof(h));
      if (!comm) return;
      const unsigned seq = comm->seq();
      const int b = seq & 1;
      for (int d = 0; d < 4; d++) {
        h.comm_dim[d] = comm->comm_dim[d] && (!comm_override || comm_override[d]);
        for (int dir = 0; dir < 2; dir++) {
          h.ghost[d][dir] = h.comm_dim[d] ? comm->recv[b][d][dir] : nullptr;
          h.wait_flag[d][dir] = h.comm_dim[d] ? comm->recv_flag[b][d][dir] : nullptr;
        }
      }
      h.seq = seq;
      h.timeout_flag = comm->timeout_flag;
    }

    // describe the faces of `in` (single-parity field holding parity `in_parity`) and where they go; advances the exchange
    static void pack_args_for(b200_pack_args &a, const ColorSpinorField &in, int in_parity, bool dagger, const int *comm_override,
                              CommContext *comm)
    {
      const unsigned seq = ++comm->seq();
      const int b = seq & 1;
      memset(&a, 0, sizeof(a));
      a.abi_version = B200_ABI_VERSION;
      a.precision = in.precision;
      for (int d = 0; d < 4; d++) {
        a.X[d] = in.X[d];
        a.comm_dim[d] = comm->comm_dim[d] && (!comm_override || comm_override[d]);
        for (int f = 0; f < 2; f++) {
          a.dst[d][f] = comm->send_dst[b][d][f];
          a.signal[d][f] = comm->send_signal[b][d][f];
        }
      }
      a.parity = in_parity;
      a.dagger = dagger;
      a.in = in.desc();
      a.block_counter = comm->block_counter;
      a.seq = seq;
    }

    // two-stream schedule: ship the faces on the side stream
    static void exchange_start(const ColorSpinorField &in, int in_parity, bool dagger, const int *comm_override,
                               CommContext *comm, void *stream)
    {
      b200_pack_args a;
      pack_args_for(a, in, in_parity, dagger, comm_override, comm);
      if (comm->pack_stream && comm->pack_stream != stream) {
        // fork: the pack kernel runs on its own stream, concurrently with the interior tiles (joined in apply())
        StreamEvents &ev = events_of(comm);
        cuda_ok(cudaEventRecord(ev.fork, cs(stream)), "record");
        cuda_ok(cudaStreamWaitEvent(cs(comm->pack_stream), ev.fork, 0), "wait");
        a.stream = comm->pack_stream;
      } else {
        a.stream = stream;
      }
      abi_ok(b200_pack_ghost(&a));
    }

    // How a partitioned Dslash is issued (B200_HALO_SCHEDULE):
    //   split  two launches of dslash_fused_kernel: [pack | boundary] on the high-priority side stream, [interior] on the
    //          operator's stream, joined by an event (measured worst: the boundary CTAs spin from the start of the launch)
    //   fused  ONE launch [pack | interior | boundary] on the operator's stream (no side stream, no events)
    //   streams  the round-1 schedule: pack kernel + boundary-tile kernel on the side stream, interior tiles on the main one
    //   sites  as streams, but with 1-site-thick shells: pack kernel, then the boundary-site role on the side stream, the
    //          interior-site role on the main stream (18.75 % instead of 28 % of the sites take the branchy path at 8 GPUs)
    enum HaloSchedule { SCHED_AUTO = -1, SCHED_SPLIT = 0, SCHED_FUSED = 1, SCHED_STREAMS = 2, SCHED_SITES = 3 };
    // Default (B200_HALO_SCHEDULE unset): `streams` -- measured best at 2 and 8 GPUs for y/z/t splits -- unless x is
    // partitioned: tiles hold full x rows, so with an x split EVERY tile is a boundary tile and the tile-granular schedule has
    // no interior left to overlap with the halo (90.6 us at 2 GPUs); there the 1-site-thick shells of `sites` are used.
    static HaloSchedule halo_schedule(const int *comm_dim_eff)
    {
      static int v = -2;
      if (v == -2) {
        const char *e = getenv("B200_HALO_SCHEDULE");
        v = SCHED_AUTO;
        if (e && strcmp(e, "split") == 0) v = SCHED_SPLIT;
        if (e && strcmp(e, "fused") == 0) v = SCHED_FUSED;
        if (e && strcmp(e, "streams") == 0) v = SCHED_STREAMS;
        if (e && strcmp(e, "sites") == 0) v = SCHED_SITES;
      }
      if (v != SCHED_AUTO) return (HaloSchedule)v;
      return comm_dim_eff[0] ? SCHED_SITES : SCHED_STREAMS;
    }

    // `b`, `asymmetric`: twisted mass only; with_x: 1 / 0 force x on / off (the twisted-mass preconditioned operator's
    // xpay flag), -1 = the Wilson convention (x iff a != 0)
    static void apply(int op, ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, const CloverField *A,
                      bool inverse_field, double a, const ColorSpinorField &x, int parity, bool dagger,
                      const int *comm_override, CommContext *comm, void *stream, double b = 0.0, bool asymmetric = false,
                      int with_x = -1)
    {
      if (in.precision != U.precision)
        throw Error("spinor precision " + std::to_string(in.precision) + " does not match the gauge field's " + std::to_string(U.precision));
      b200_dslash_args args;
      memset(&args, 0, sizeof(args));
      args.abi_version = B200_ABI_VERSION;
      args.op = op;
      args.kernel = B200_KERNEL_AUTO;
      args.precision = in.precision;
      for (int d = 0; d < 4; d++) args.X[d] = U.X[d];
      args.parity = parity == QUDA_INVALID_PARITY ? 0 : parity;
      args.dagger = dagger;
      args.a = a;
      args.out = out.desc();
      args.in = in.desc();
      if (with_x < 0 ? a != 0.0 : with_x != 0) args.x = x.desc();
      args.b = b;
      args.asymmetric = asymmetric ? 1 : 0;
      args.U = U.g;
      if (A) args.A = (inverse_field && A->has_inverse()) ? A->cinv : A->c;
      bool part = false;
      if (comm)
        for (int d = 0; d < 4; d++) part |= (comm->comm_dim[d] && (!comm_override || comm_override[d]));
      if (part && in.n_parity != 1) throw Error("a partitioned Dslash works on one parity at a time");
      const bool side_stream = part && comm->pack_stream && comm->pack_stream != stream;
      int eff[4] = {0, 0, 0, 0};
      if (part)
        for (int d = 0; d < 4; d++) eff[d] = comm->comm_dim[d] && (!comm_override || comm_override[d]);
      const HaloSchedule sched = part ? halo_schedule(eff) : SCHED_STREAMS;
      if (part && (sched == SCHED_FUSED || (sched == SCHED_SPLIT && !side_stream))) {
        // pack + interior + boundary: one launch on the operator's stream
        b200_pack_args pk;
        pack_args_for(pk, in, 1 - parity, dagger, comm_override, comm);
        pk.stream = stream;
        halo_fill(args.halo, comm_override, comm);
        args.stream = stream;
        abi_ok(b200_dslash_apply_fused(&args, &pk));
        return;
      }
      if (part && sched == SCHED_SPLIT) {
        // [pack | boundary] on the side stream (after `in` is complete), [interior] on the main stream, then join
        b200_pack_args pk;
        pack_args_for(pk, in, 1 - parity, dagger, comm_override, comm);
        halo_fill(args.halo, comm_override, comm);
        StreamEvents &ev = events_of(comm);
        cuda_ok(cudaEventRecord(ev.fork, cs(stream)), "record");
        cuda_ok(cudaStreamWaitEvent(cs(comm->pack_stream), ev.fork, 0), "wait");
        pk.stream = comm->pack_stream;
        args.stream = comm->pack_stream;
        args.kernel = B200_KERNEL_BOUNDARY_TILES;
        abi_ok(b200_dslash_apply_fused(&args, &pk));
        pk.stream = stream;
        args.stream = stream;
        args.kernel = B200_KERNEL_INTERIOR_TILES;
        abi_ok(b200_dslash_apply_fused(&args, &pk));
        cuda_ok(cudaEventRecord(ev.join, cs(comm->pack_stream)), "record");
        cuda_ok(cudaStreamWaitEvent(cs(stream), ev.join, 0), "wait");
        return;
      }
      if (part) exchange_start(in, 1 - parity, dagger, comm_override, comm, stream);
      halo_fill(args.halo, comm_override, part ? comm : nullptr);
      const bool two_streams = part && comm->pack_stream && comm->pack_stream != stream;
      if (two_streams) {
        // side stream (behind the pack kernel): boundary sites -- they depend only on the halo, not on the interior
        // launch; main stream: the interior.  Both halves write disjoint sites.
        const bool shells = sched == SCHED_SITES;
        args.kernel = shells ? B200_KERNEL_BOUNDARY_SITES : B200_KERNEL_BOUNDARY_TILES;
        args.stream = comm->pack_stream;
        abi_ok(b200_dslash_apply(&args));
        args.kernel = shells ? B200_KERNEL_INTERIOR_SITES : B200_KERNEL_INTERIOR_TILES;
        args.stream = stream;
        abi_ok(b200_dslash_apply(&args));
        // join: `out` is complete, and `in` may be overwritten, only after the side stream has drained
        StreamEvents &ev = events_of(comm);
        cuda_ok(cudaEventRecord(ev.join, cs(comm->pack_stream)), "record");
        cuda_ok(cudaStreamWaitEvent(cs(stream), ev.join, 0), "wait");
      } else {
        args.stream = stream;
        abi_ok(b200_dslash_apply(&args));
      }
    }

    void ApplyWilson(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, double a,
                     const ColorSpinorField &x, int parity, bool dagger, const int *comm_override, CommContext *comm,
                     void *stream)
    {
      apply(B200_OP_WILSON, out, in, U, nullptr, false, a, x, parity, dagger, comm_override, comm, stream);
    }

    void ApplyWilsonClover(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, const CloverField &A,
                           double a, const ColorSpinorField &x, int parity, bool dagger, const int *comm_override,
                           CommContext *comm, void *stream)
    {
      apply(B200_OP_CLOVER, out, in, U, &A, false, a, x, parity, dagger, comm_override, comm, stream);
    }

    void ApplyWilsonCloverPreconditioned(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U,
                                         const CloverField &A, double a, const ColorSpinorField &x, int parity,
                                         bool dagger, const int *comm_override, CommContext *comm, void *stream)
    {
      apply(B200_OP_CLOVER_PC, out, in, U, &A, true, a, x, parity, dagger, comm_override, comm, stream);
    }

    void ApplyTwistedMass(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, double a, double b,
                          const ColorSpinorField &x, int parity, bool dagger, const int *comm_override, CommContext *comm,
                          void *stream)
    {
      apply(B200_OP_TWISTED_MASS, out, in, U, nullptr, false, a, x, parity, dagger, comm_override, comm, stream, b);
    }

    void ApplyTwistedMassPreconditioned(ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, double a,
                                        double b, bool xpay, const ColorSpinorField &x, int parity, bool dagger,
                                        bool asymmetric, const int *comm_override, CommContext *comm, void *stream)
    {
      apply(B200_OP_TWISTED_MASS_PC, out, in, U, nullptr, false, a, x, parity, dagger, comm_override, comm, stream, b,
            asymmetric, xpay ? 1 : 0);
    }

    void ApplyTwistGamma(ColorSpinorField &out, const ColorSpinorField &in, double kappa, double mu, bool dagger, bool inverse,
                         void *stream)
    {
      b200_spinor o = out.desc(), i = in.desc();
      abi_ok(b200_twist_gamma5(&o, &i, in.precision, kappa, mu, dagger, inverse, stream));
    }

    void ApplyClover(ColorSpinorField &out, const ColorSpinorField &in, const CloverField &A, bool inverse, int parity,
                     void *stream)
    {
      b200_spinor o = out.desc(), i = in.desc();
      b200_clover c = (inverse && A.has_inverse()) ? A.cinv : A.c;
      abi_ok(b200_clover_apply(&o, &i, &c, in.precision, inverse, parity, stream));
    }

    bool halo_timed_out(CommContext *comm, void *stream)
    {
      if (!comm || !comm->timeout_flag) return false;
      int flag = 0;
      cuda_ok(cudaMemcpyAsync(&flag, comm->timeout_flag, sizeof(int), cudaMemcpyDeviceToHost, cs(stream)), "memcpy(timeout flag)");
      cuda_ok(cudaStreamSynchronize(cs(stream)), "sync");
      if (flag) cuda_ok(cudaMemsetAsync(comm->timeout_flag, 0, sizeof(int), cs(stream)), "memset(timeout flag)");
      return flag != 0;
    }

    // ------------------------------------------------------------------ blas + reductions
    namespace blas
    {
      static long long g_flops = 0;
      long long flops() { return g_flops; }

      constexpr int kBlocks = 148 * 4, kThreads = 256, kMaxVals = 2;

      // device-resident scalars of a solve (read by the update kernels, written by the reduction finalisers)
      enum Scalar { S_R2 = 0, S_R2_OLD = 1, S_PAP = 2, S_ALPHA = 3, S_BETA = 4, S_RAW0 = 5, S_RAW1 = 6, S_COUNT = 8 };
      enum Finish { FIN_RAW = 0, FIN_PAP = 1, FIN_R2 = 2 };

      // NVLink mailboxes of the all-reduce (b200_comm::reduce_peer): slot [parity of seq][source rank] in every rank's box
      struct ReduceSlot {
        double v[4];
        unsigned seq;
        unsigned pad[7];
      };
      static_assert(sizeof(ReduceSlot) == B200_REDUCE_SLOT_BYTES, "mailbox slot layout");
      struct ReducePeers {
        ReduceSlot *box[B200_MAX_RANKS];
        int rank, n_ranks;
        unsigned seq;
        int *timeout_flag;
      };

      // per-stream reduction workspace
      struct Workspace {
        double *partials = nullptr; // [kBlocks][kMaxVals]
        unsigned *ticket = nullptr;
        double *scalars = nullptr;  // [S_COUNT] on the device
        double *host = nullptr;     // pinned + mapped: [ring][S_COUNT], written by the finalisers
        double *host_dev = nullptr; // device alias of `host`
        cudaEvent_t ev[8] = {};
        unsigned long long ring = 0;
      };
      static Workspace &workspace(void *stream)
      {
        static std::map<void *, Workspace> m;
        Workspace &w = m[stream];
        if (!w.partials) {
          cuda_ok(cudaMalloc(&w.partials, sizeof(double) * kBlocks * kMaxVals), "cudaMalloc(reduce)");
          cuda_ok(cudaMalloc(&w.ticket, sizeof(unsigned)), "cudaMalloc(reduce)");
          cuda_ok(cudaMemset(w.ticket, 0, sizeof(unsigned)), "memset");
          cuda_ok(cudaMalloc(&w.scalars, sizeof(double) * S_COUNT), "cudaMalloc(reduce)");
          cuda_ok(cudaMemset(w.scalars, 0, sizeof(double) * S_COUNT), "memset");
          cuda_ok(cudaHostAlloc(&w.host, sizeof(double) * 8 * S_COUNT, cudaHostAllocMapped), "cudaHostAlloc(reduce)");
          cuda_ok(cudaHostGetDevicePointer(&w.host_dev, w.host, 0), "cudaHostGetDevicePointer");
          for (auto &e : w.ev) cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
        }
        return w;
      }

      // Second stage of every reduction, run by the block that arrives last: sum the per-block partial sums in block
      // order (fixed -> bit-reproducible), all-reduce over the ranks through the NVLink mailboxes in rank order, derive
      // the CG scalars and publish everything to the host mirror.
      template <int NV>
      __device__ void finish_reduction(const double *acc_block, double *partials, unsigned *ticket, double *S, double *host_out,
                                       int fin, ReducePeers peers)
      {
        __shared__ double sh[kThreads][kMaxVals];
        __shared__ bool is_last;
        __shared__ double part[B200_MAX_RANKS][kMaxVals];
        const int t = threadIdx.x;
        if (t == 0) {
          for (int i = 0; i < NV; i++) partials[blockIdx.x * kMaxVals + i] = acc_block[i];
          __threadfence();
          is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        double a[kMaxVals] = {0, 0};
        for (int b = t; b < (int)gridDim.x; b += kThreads)
          for (int i = 0; i < NV; i++) a[i] += __ldcg(partials + b * kMaxVals + i);
        for (int i = 0; i < NV; i++) sh[t][i] = a[i];
        __syncthreads();
        for (int s = kThreads / 2; s > 0; s >>= 1) {
          if (t < s)
            for (int i = 0; i < NV; i++) sh[t][i] += sh[t + s][i];
          __syncthreads();
        }
        if (peers.n_ranks > 1) {
          const int b = peers.seq & 1;
          if (t < peers.n_ranks) {
            ReduceSlot *dst = peers.box[t] + b * B200_MAX_RANKS + peers.rank;
            for (int i = 0; i < NV; i++) dst->v[i] = sh[0][i];
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&dst->seq), "r"(peers.seq) : "memory");
            const ReduceSlot *src = peers.box[peers.rank] + b * B200_MAX_RANKS + t;
            const long long t0 = clock64();
            for (;;) {
              unsigned got;
              asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(&src->seq) : "memory");
              if ((int)(got - peers.seq) >= 0) break;
              if (clock64() - t0 > 20000000000LL) { // ~10 s: a lost peer must never hang the GPU
                if (peers.timeout_flag) *peers.timeout_flag = 1;
                break;
              }
              __nanosleep(40);
            }
            for (int i = 0; i < NV; i++) part[t][i] = *reinterpret_cast<const volatile double *>(&src->v[i]);
          }
          __syncthreads();
          if (t == 0)
            for (int i = 0; i < NV; i++) {
              double s = 0;
              for (int r = 0; r < peers.n_ranks; r++) s += part[r][i];
              sh[0][i] = s;
            }
        }
        if (t == 0) {
          const double v0 = sh[0][0];
          if (fin == FIN_PAP) {
            S[S_PAP] = v0;
            S[S_ALPHA] = S[S_R2] / v0;
          } else if (fin == FIN_R2) {
            const double old = S[S_R2];
            S[S_R2_OLD] = old;
            S[S_R2] = v0;
            S[S_BETA] = v0 / old;
          }
          S[S_RAW0] = v0;
          if (NV > 1) S[S_RAW1] = sh[0][1];
          if (host_out) {
            for (int i = 0; i < S_COUNT; i++) host_out[i] = S[i];
            __threadfence_system();
          }
          *ticket = 0;
        }
      }

      __device__ __forceinline__ double warp_sum(double v)
      {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        return v;
      }
      // per-block sum of per-thread values, in a fixed order (warp tree, then warps in index order)
      __device__ __forceinline__ double block_sum(double v, double *warp_buf)
      {
        v = warp_sum(v);
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) warp_buf[w] = v;
        __syncthreads();
        double s = 0;
        if (threadIdx.x == 0)
          for (int i = 0; i < kThreads / 32; i++) s += warp_buf[i];
        __syncthreads();
        return s;
      }

      struct ReduceArgs {
        double *partials;
        unsigned *ticket;
        double *S;
        double *host_out;
        int fin;
        ReducePeers peers;
      };

      // y = a x + b y with an optional reduction over the result (R_NORM_Y) or <x, y> (R_DOT_XY)
      enum { R_NONE = 0, R_NORM_Y = 1, R_DOT_XY = 2 };
      template <typename Tx, typename Ty, int R>
      __global__ void __launch_bounds__(kThreads) axpby_kernel(double a, const Tx *__restrict__ x, double b, Ty *__restrict__ y,
                                                               size_t n, bool write, ReduceArgs ra)
      {
        double acc = 0;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
          const double xv = x[i], yv = y[i];
          const double r = a * xv + b * yv;
          if (write) y[i] = (Ty)r;
          if (R == R_NORM_Y) acc += (write ? (double)(Ty)r * (double)(Ty)r : yv * yv);
          if (R == R_DOT_XY) acc += xv * yv;
        }
        if (R != R_NONE) {
          __shared__ double wb[kThreads / 32];
          const double s = block_sum(acc, wb);
          finish_reduction<1>(&s, ra.partials, ra.ticket, ra.S, ra.host_out, ra.fin, ra.peers);
        }
      }

      // ---- the three kernels of a CG iteration; alpha and beta come from the device scalars
      // r -= alpha Ap ; |r|^2   (finaliser: r2_old <- r2, r2 <- |r|^2, beta <- r2 / r2_old)
      template <typename T>
      __global__ void __launch_bounds__(kThreads) cg_update_r_kernel(T *__restrict__ r, const T *__restrict__ Ap, size_t n, ReduceArgs ra)
      {
        const double alpha = ra.S[S_ALPHA];
        double acc = 0;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
          const T v = (T)((double)r[i] - alpha * (double)Ap[i]);
          r[i] = v;
          acc += (double)v * (double)v;
        }
        __shared__ double wb[kThreads / 32];
        const double s = block_sum(acc, wb);
        finish_reduction<1>(&s, ra.partials, ra.ticket, ra.S, ra.host_out, ra.fin, ra.peers);
      }
      // x += alpha p ; p = r + beta p   (the reference's axpyZpbx, lib/inv_cg_quda.cpp:389)
      template <typename T>
      __global__ void __launch_bounds__(kThreads) cg_update_xp_kernel(T *__restrict__ x, T *__restrict__ p, const T *__restrict__ r,
                                                                      size_t n, const double *__restrict__ S)
      {
        const double alpha = S[S_ALPHA], beta = S[S_BETA];
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
          const double pv = p[i];
          x[i] = (T)((double)x[i] + alpha * pv);
          p[i] = (T)((double)r[i] + beta * pv);
        }
      }
      // after a reliable update: p += r_new - r_old ; r_old <- r_new   (keeps p = r + beta p_old with the true residual)
      template <typename T>
      __global__ void __launch_bounds__(kThreads) cg_replace_r_kernel(T *__restrict__ p, T *__restrict__ r, const T *__restrict__ r_new, size_t n)
      {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
          const T rn = r_new[i];
          p[i] = (T)((double)p[i] + ((double)rn - (double)r[i]));
          r[i] = rn;
        }
      }
      __global__ void set_scalar_kernel(double *S, int idx, double v) { S[idx] = v; }

      static void check_pair(const ColorSpinorField &x, const ColorSpinorField &y)
      {
        if (x.Length() != y.Length()) throw Error("blas: operands differ in length");
        if (x.precision == B200_HALF || y.precision == B200_HALF) throw Error("blas: half-precision (block-float) fields are not supported");
      }

      static ReducePeers peers_of(CommContext *comm)
      {
        ReducePeers p;
        memset(&p, 0, sizeof(p));
        p.n_ranks = 1;
        if (comm && comm->mailboxes()) {
          if (comm->n_ranks > B200_MAX_RANKS) throw Error("NVLink all-reduce: more ranks than mailbox slots");
          for (int r = 0; r < B200_MAX_RANKS; r++) p.box[r] = reinterpret_cast<ReduceSlot *>(comm->reduce_peer[r]);
          p.rank = comm->rank;
          p.n_ranks = comm->n_ranks;
          p.seq = ++comm->reduce_seq();
          p.timeout_flag = comm->timeout_flag;
        }
        return p;
      }

      // reduction launch bookkeeping: where the finaliser writes, and the event the host may wait on
      struct Pending {
        Workspace *w;
        int slot;
      };
      static ReduceArgs reduce_args(const Exec &ex, int fin, Pending &pend)
      {
        Workspace &w = workspace(ex.stream);
        pend.w = &w;
        pend.slot = (int)(w.ring++ & 7);
        ReduceArgs ra;
        ra.partials = w.partials;
        ra.ticket = w.ticket;
        ra.S = w.scalars;
        ra.host_out = w.host_dev + pend.slot * S_COUNT;
        ra.fin = fin;
        ra.peers = peers_of(ex.comm);
        return ra;
      }
      static void mark(const Exec &ex, const Pending &p) { cuda_ok(cudaEventRecord(p.w->ev[p.slot], cs(ex.stream)), "record"); }
      // wait for a reduction launched earlier and return the host mirror of the device scalars as of that launch;
      // multi-rank sums without NVLink mailboxes go through the host callback here
      static const double *await(const Exec &ex, const Pending &p, bool need_host_allreduce, double *scratch)
      {
        cuda_ok(cudaEventSynchronize(p.w->ev[p.slot]), "event sync");
        const double *h = p.w->host + p.slot * S_COUNT;
        if (!need_host_allreduce) return h;
        memcpy(scratch, h, sizeof(double) * S_COUNT);
        ex.comm->allreduce_sum(&scratch[S_RAW0], 1, ex.comm->user);
        return scratch;
      }
      static bool host_allreduce(const Exec &ex) { return ex.comm && !ex.comm->mailboxes() && ex.comm->allreduce_sum; }

      template <int R> static double run(double a, const ColorSpinorField &x, double b, ColorSpinorField &y, bool write, const Exec &ex)
      {
        check_pair(x, y);
        if (x.precision != y.precision) throw Error("blas: operands differ in precision (convert with blas::copy)");
        const size_t n = x.Length();
        Pending pend {};
        ReduceArgs ra {};
        if (R != R_NONE) ra = reduce_args(ex, FIN_RAW, pend);
        cudaStream_t s = cs(ex.stream);
        if (x.precision == 8)
          axpby_kernel<double, double, R><<<kBlocks, kThreads, 0, s>>>(a, (const double *)x.v, b, (double *)y.v, n, write, ra);
        else
          axpby_kernel<float, float, R><<<kBlocks, kThreads, 0, s>>>(a, (const float *)x.v, b, (float *)y.v, n, write, ra);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 3 * (long long)n;
        if (R == R_NONE) return 0.0;
        mark(ex, pend);
        double scratch[S_COUNT];
        return await(ex, pend, host_allreduce(ex), scratch)[S_RAW0];
      }

      // Precision conversion has to go through the site/component map: the native order of fp64 fields is planes of 2
      // reals, that of fp32 fields planes of 4 (color_spinor_field_order.h FloatNOrder), so an element-wise cast
      // would permute components.  One thread per site moves its 24 reals.
      template <typename Ts, int Ns, typename Td, int Nd>
      __global__ void convert_kernel(const Ts *__restrict__ src, Td *__restrict__ dst, int volume_cb, size_t src_parity_elems,
                                     size_t dst_parity_elems)
      {
        const int x = blockIdx.x * blockDim.x + threadIdx.x;
        if (x >= volume_cb) return;
        const Ts *s = src + blockIdx.y * src_parity_elems;
        Td *d = dst + blockIdx.y * dst_parity_elems;
        double v[24];
#pragma unroll
        for (int r = 0; r < 24; r++) v[r] = s[((size_t)(r / Ns) * volume_cb + x) * Ns + r % Ns];
#pragma unroll
        for (int r = 0; r < 24; r++) d[((size_t)(r / Nd) * volume_cb + x) * Nd + r % Nd] = (Td)v[r];
      }

      void copy(ColorSpinorField &dst, const ColorSpinorField &src, const Exec &ex)
      {
        check_pair(src, dst);
        if (dst.n_parity != src.n_parity) throw Error("blas::copy: fields cover different site subsets");
        if (src.precision == dst.precision) {
          if (dst.v != src.v) cuda_ok(cudaMemcpyAsync(dst.v, src.v, src.Bytes(), cudaMemcpyDeviceToDevice, cs(ex.stream)), "copy");
          return;
        }
        const int vcb = src.VolumeCB();
        dim3 grid((vcb + 127) / 128, src.n_parity);
        const size_t se = (size_t)24 * vcb, de = (size_t)24 * vcb;
        if (src.precision == 8)
          convert_kernel<double, 2, float, 4><<<grid, 128, 0, cs(ex.stream)>>>((const double *)src.v, (float *)dst.v, vcb, se, de);
        else
          convert_kernel<float, 4, double, 2><<<grid, 128, 0, cs(ex.stream)>>>((const float *)src.v, (double *)dst.v, vcb, se, de);
        cuda_ok(cudaGetLastError(), "convert launch");
      }
      void zero(ColorSpinorField &x, const Exec &ex) { cuda_ok(cudaMemsetAsync(x.v, 0, x.Bytes(), cs(ex.stream)), "memset"); }
      void ax(double a, ColorSpinorField &x, const Exec &ex) { run<R_NONE>(0.0, x, a, x, true, ex); }
      void axpy(double a, const ColorSpinorField &x, ColorSpinorField &y, const Exec &ex) { run<R_NONE>(a, x, 1.0, y, true, ex); }
      void xpay(const ColorSpinorField &x, double a, ColorSpinorField &y, const Exec &ex) { run<R_NONE>(1.0, x, a, y, true, ex); }
      void axpby(double a, const ColorSpinorField &x, double b, ColorSpinorField &y, const Exec &ex) { run<R_NONE>(a, x, b, y, true, ex); }
      double norm2(const ColorSpinorField &x, const Exec &ex)
      {
        return run<R_NORM_Y>(0.0, x, 1.0, const_cast<ColorSpinorField &>(x), false, ex);
      }
      double reDotProduct(const ColorSpinorField &x, const ColorSpinorField &y, const Exec &ex)
      {
        return run<R_DOT_XY>(0.0, x, 1.0, const_cast<ColorSpinorField &>(y), false, ex);
      }
      double axpyNorm(double a, const ColorSpinorField &x, ColorSpinorField &y, const Exec &ex) { return run<R_NORM_Y>(a, x, 1.0, y, true, ex); }
      double xmyNorm(const ColorSpinorField &x, ColorSpinorField &y, const Exec &ex) { return run<R_NORM_Y>(1.0, x, -1.0, y, true, ex); }

      // ---- CG iteration pieces (used by invertCG below)
      static void set_scalar(const Exec &ex, int idx, double v)
      {
        set_scalar_kernel<<<1, 1, 0, cs(ex.stream)>>>(workspace(ex.stream).scalars, idx, v);
        cuda_ok(cudaGetLastError(), "scalar launch");
      }
      // <p, Ap> -> pAp, alpha = r2 / pAp (device)
      static Pending cg_dot(const ColorSpinorField &p, const ColorSpinorField &Ap, const Exec &ex)
      {
        check_pair(p, Ap);
        Pending pend {};
        ReduceArgs ra = reduce_args(ex, FIN_PAP, pend);
        const size_t n = p.Length();
        if (p.precision == 8)
          axpby_kernel<double, double, R_DOT_XY><<<kBlocks, kThreads, 0, cs(ex.stream)>>>(0.0, (const double *)p.v, 1.0, (double *)Ap.v, n, false, ra);
        else
          axpby_kernel<float, float, R_DOT_XY><<<kBlocks, kThreads, 0, cs(ex.stream)>>>(0.0, (const float *)p.v, 1.0, (float *)Ap.v, n, false, ra);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 2 * (long long)n;
        mark(ex, pend);
        return pend;
      }
      static Pending cg_update_r(ColorSpinorField &r, const ColorSpinorField &Ap, const Exec &ex)
      {
        check_pair(r, Ap);
        Pending pend {};
        ReduceArgs ra = reduce_args(ex, FIN_R2, pend);
        const size_t n = r.Length();
        if (r.precision == 8)
          cg_update_r_kernel<double><<<kBlocks, kThreads, 0, cs(ex.stream)>>>((double *)r.v, (const double *)Ap.v, n, ra);
        else
          cg_update_r_kernel<float><<<kBlocks, kThreads, 0, cs(ex.stream)>>>((float *)r.v, (const float *)Ap.v, n, ra);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 4 * (long long)n;
        mark(ex, pend);
        return pend;
      }
      static void cg_update_xp(ColorSpinorField &x, ColorSpinorField &p, const ColorSpinorField &r, const Exec &ex)
      {
        check_pair(x, p);
        const size_t n = p.Length();
        const double *S = workspace(ex.stream).scalars;
        if (p.precision == 8)
          cg_update_xp_kernel<double><<<kBlocks, kThreads, 0, cs(ex.stream)>>>((double *)x.v, (double *)p.v, (const double *)r.v, n, S);
        else
          cg_update_xp_kernel<float><<<kBlocks, kThreads, 0, cs(ex.stream)>>>((float *)x.v, (float *)p.v, (const float *)r.v, n, S);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 4 * (long long)n;
      }
      static void cg_replace_r(ColorSpinorField &p, ColorSpinorField &r, const ColorSpinorField &r_new, const Exec &ex)
      {
        check_pair(p, r_new);
        const size_t n = p.Length();
        if (p.precision == 8)
          cg_replace_r_kernel<double><<<kBlocks, kThreads, 0, cs(ex.stream)>>>((double *)p.v, (double *)r.v, (const double *)r_new.v, n);
        else
          cg_replace_r_kernel<float><<<kBlocks, kThreads, 0, cs(ex.stream)>>>((float *)p.v, (float *)r.v, (const float *)r_new.v, n);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 2 * (long long)n;
      }
    } // namespace blas

    // ------------------------------------------------------------------ the even-odd operator
    Dirac::Dirac(SiteTerm term_, bool schur_, const DiracParam &p) :
      gauge(p.gauge), clover(p.clover), kappa(p.kappa), mu(p.mu), matpcType(p.matpcType), dagger(p.dagger), comm(p.comm),
      stream(p.stream), term(term_), schur(schur_)
    {
      if (!gauge) throw Error("operator needs a gauge field");
      if (term == SiteTerm::Clover) {
        if (!clover) throw Error("clover operator needs a clover field");
        if (schur && !clover->has_inverse() && !clover->c.dynamic_inverse)
          throw Error("even-odd preconditioned clover operator needs A^-1 (a static inverse field or dynamic_inverse)");
      }
      for (int d = 0; d < 4; d++) commDim[d] = p.commDim[d];
      symmetric = (matpcType == QUDA_MATPC_EVEN_EVEN || matpcType == QUDA_MATPC_ODD_ODD);
      this_parity = (matpcType == QUDA_MATPC_EVEN_EVEN || matpcType == QUDA_MATPC_EVEN_EVEN_ASYMMETRIC) ? 0 : 1;
      other_parity = 1 - this_parity;
    }

    Dirac *Dirac::create(const std::string &type, const DiracParam &p)
    {
      if (type == "wilson") return new Dirac(SiteTerm::Identity, false, p);
      if (type == "wilsonpc") return new Dirac(SiteTerm::Identity, true, p);
      if (type == "clover") return new Dirac(SiteTerm::Clover, false, p);
      if (type == "cloverpc") return new Dirac(SiteTerm::Clover, true, p);
      if (type == "twistedmass") return new Dirac(SiteTerm::Twist, false, p);
      if (type == "twistedmasspc") return new Dirac(SiteTerm::Twist, true, p);
      throw Error("no operator of type '" + type + "' in this engine");
    }

    static void need_single_parity(const ColorSpinorField &a, const ColorSpinorField &b)
    {
      if (a.n_parity != 1 || b.n_parity != 1) throw Error("this operation acts on single-parity fields");
      if (a.v == b.v) throw Error("input and output must be different fields");
    }
    static void need_full(const ColorSpinorField &a, const ColorSpinorField &b)
    {
      if (a.n_parity != 2 || b.n_parity != 2) throw Error("this operation acts on full (two-parity) fields");
    }

    // One Dslash launch: out = [x +] k * (site-term fused in the epilogue) D in
    //   Fuse::None     : out = D in                | x + k D in
    //   Fuse::A        : out = A x + k D in        (always with x)
    //   Fuse::AinvPost : out = A^-1 D in           | x + k A^-1 D in
    // For the identity site term all three coincide with the Wilson form.
    void Dirac::hop(ColorSpinorField &out, const ColorSpinorField &in, int parity, Fuse f, const ColorSpinorField *x, double k) const
    {
      need_single_parity(in, out);
      const ColorSpinorField &xf = x ? *x : in;
      const double a = x ? k : 0.0;
      if (term == SiteTerm::Identity || f == Fuse::None) {
        ApplyWilson(out, in, *gauge, a, xf, parity, dagger, commDim, comm, stream);
      } else if (term == SiteTerm::Clover) {
        if (f == Fuse::A) {
          if (!x) throw Error("A x + k D in needs x");
          ApplyWilsonClover(out, in, *gauge, *clover, a, xf, parity, dagger, commDim, comm, stream);
        } else {
          ApplyWilsonCloverPreconditioned(out, in, *gauge, *clover, a, xf, parity, dagger, commDim, comm, stream);
        }
      } else { // twist: A = 1 + i 2 kappa mu gamma5, A^-1 = (1 - i 2 kappa mu gamma5) / (1 + (2 kappa mu)^2)
        if (f == Fuse::A) {
          if (!x) throw Error("the twisted-mass Dslash exists only in its xpay form");
          ApplyTwistedMass(out, in, *gauge, a, 2 * mu * kappa, xf, parity, dagger, commDim, comm, stream);
        } else {
          const double tw = -2.0 * kappa * mu;
          const double nrm = 1.0 / (1.0 + tw * tw);
          const bool asym = !symmetric && dagger;
          ApplyTwistedMassPreconditioned(out, in, *gauge, x ? k * nrm : nrm, tw, x != nullptr, xf, parity, dagger, asym, commDim, comm, stream);
        }
      }
      dslash_applications++;
    }

    // out = A in or A^-1 in on one parity
    void Dirac::site(ColorSpinorField &out, const ColorSpinorField &in, int parity, bool inverse) const
    {
      switch (term) {
      case SiteTerm::Identity: blas::copy(out, in, exec()); break;
      case SiteTerm::Clover: ApplyClover(out, in, *clover, inverse, parity, stream); break;
      case SiteTerm::Twist: ApplyTwistGamma(out, in, kappa, mu, dagger, inverse, stream); break;
      }
    }

    void Dirac::Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const
    {
      if (term == SiteTerm::Twist && !schur) throw Error("the unpreconditioned twisted-mass Dslash exists only in its xpay form");
      hop(out, in, parity, schur ? Fuse::AinvPost : Fuse::None, nullptr, 0.0);
    }

    void Dirac::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x, double k) const
    {
      hop(out, in, parity, schur ? Fuse::AinvPost : Fuse::A, &x, k);
    }

    void Dirac::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      if (!schur) {
        // out_p = A_p in_p - kappa D in_{1-p}
        need_full(out, in);
        const bool one_launch = term != SiteTerm::Twist && !(comm && comm->partitioned());
        if (one_launch) { // both parities in one launch (full-field kernel)
          if (term == SiteTerm::Identity)
            ApplyWilson(out, in, *gauge, -kappa, in, QUDA_INVALID_PARITY, dagger, commDim, comm, stream);
          else
            ApplyWilsonClover(out, in, *gauge, *clover, -kappa, in, QUDA_INVALID_PARITY, dagger, commDim, comm, stream);
          dslash_applications += 2;
        } else {
          for (int p = 0; p < 2; p++) {
            auto o = out.parity_view(p);
            const auto x = in.parity_view(p);
            hop(o, in.parity_view(1 - p), p, Fuse::A, &x, -kappa);
          }
        }
        return;
      }
      // Schur complement on `this_parity`
      const double k2 = -kappa * kappa;
      Scratch tmp(stream, in, 1);
      if (term == SiteTerm::Clover && symmetric && dagger) {
        // (1 - k^2 A^-1 D A^-1 D)^dagger = 1 - k^2 D^dagger A^-1 D^dagger A^-1   (A hermitian)
        site(out, in, this_parity, true);
        hop(tmp, out, other_parity, Fuse::AinvPost, nullptr, 0.0);
        hop(out, tmp, this_parity, Fuse::None, &in, k2);
      } else {
        hop(tmp, in, other_parity, Fuse::AinvPost, nullptr, 0.0);
        hop(out, tmp, this_parity, symmetric ? Fuse::AinvPost : Fuse::A, &in, k2);
      }
    }

    void Dirac::Mdag(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      flipDagger();
      try {
        M(out, in);
      } catch (...) {
        flipDagger();
        throw;
      }
      flipDagger();
    }

    void Dirac::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      Scratch tmp(stream, in, in.n_parity);
      M(tmp, in);
      Mdag(out, tmp);
    }

    // Full-system solve through the Schur complement: M x = b  <=>  M_pc x_e = src, x_o from x_e.
    //   symmetric : src = A_e^-1 (b_e + k D_eo A_o^-1 b_o)       asymmetric : src = b_e + k D_eo A_o^-1 b_o
    // The preconditioned source is built in the other-parity half of x, the solution lives in the this-parity half.
    void Dirac::prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                        QudaSolutionType st) const
    {
      const bool pc_solution = (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION);
      if (!schur) {
        if (pc_solution) throw Error("a preconditioned solution type needs an even-odd preconditioned operator");
        src = b;
        sol = x;
        return;
      }
      if (pc_solution) {
        src = b;
        sol = x;
        return;
      }
      need_full(x, b);
      src = x.parity_view(other_parity);
      sol = x.parity_view(this_parity);
      const auto b_this = b.parity_view(this_parity), b_other = b.parity_view(other_parity);
      if (term == SiteTerm::Identity) {
        hop(src, b_other, this_parity, Fuse::None, &b_this, kappa);
        return;
      }
      Scratch tmp(stream, b, 1);
      if (symmetric) {
        site(src, b_other, other_parity, true);
        hop(tmp, src, this_parity, Fuse::None, &b_this, kappa);
        site(src, tmp, this_parity, true);
      } else {
        site(tmp, b_other, other_parity, true);
        hop(src, tmp, this_parity, Fuse::None, &b_this, kappa);
      }
    }

    // x_o = A_o^-1 (b_o + k D_oe x_e)
    void Dirac::reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType st) const
    {
      if (!schur || st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) return;
      need_full(x, b);
      auto x_other = x.parity_view(other_parity);
      const auto x_this = x.parity_view(this_parity), b_other = b.parity_view(other_parity);
      if (term == SiteTerm::Identity) {
        hop(x_other, x_this, other_parity, Fuse::None, &b_other, kappa);
        return;
      }
      Scratch tmp(stream, b, 1);
      hop(tmp, x_this, other_parity, Fuse::None, &b_other, kappa);
      site(x_other, tmp, other_parity, true);
    }

    // ------------------------------------------------------------------ CG (normal equations) with reliable updates
    // Same recurrences and reliable-update criterion as the reference's CG (lib/inv_cg_quda.cpp:237-420), restructured so
    // that no iteration waits for the host: pAp, r2, alpha, beta live on the device (written by the reduction
    // finalisers, read by the next update kernel); the host reads r2 of iteration k-1 while the GPU runs iteration k and
    // takes its convergence / reliable-update decisions one iteration late.  A late reliable update repairs the search
    // direction with p += r_true - r_sloppy, which restores p = r_true + beta p_old exactly.
    void invertCG(const Dirac &mat, const Dirac &matSloppy, ColorSpinorField &x, const ColorSpinorField &b, SolverParam &param)
    {
      using namespace blas;
      const Exec ex = mat.exec();
      if (matSloppy.Stream() != mat.Stream()) throw Error("precise and sloppy operators must share a stream");
      const auto t0 = std::chrono::steady_clock::now();
      const long long flops0 = blas::flops();
      const long long ds0 = mat.DslashApplications() + matSloppy.DslashApplications();
      const bool mixed = (&mat != &matSloppy);
      const int sp = matSloppy.Precision();
      if (x.precision != mat.Precision() || b.precision != mat.Precision()) throw Error("x and b must have the precise operator's precision");
      if (sp != 8 && sp != 4) throw Error("the sloppy operator must be double or single precision");
      if (mixed && sp > x.precision) throw Error("the sloppy operator is more precise than the precise one");
      const bool host_ar = host_allreduce(ex);
      int syncs = 0;

      // work fields come from the per-stream scratch pool: a second solve on the same stream allocates nothing
      Scratch r_s(ex.stream, x.X, x.precision, x.n_parity), y_s(ex.stream, x.X, x.precision, x.n_parity),
        tmp_s(ex.stream, x.X, x.precision, x.n_parity);
      ColorSpinorField &r = r_s.f;     // high-precision residual
      ColorSpinorField &y = y_s.f;     // high-precision accumulated solution
      ColorSpinorField &tmp = tmp_s.f;
      const bool same_prec = sp == x.precision;
      Scratch xS_s(ex.stream, x.X, sp, x.n_parity), p_s(ex.stream, x.X, sp, x.n_parity), Ap_s(ex.stream, x.X, sp, x.n_parity);
      std::unique_ptr<Scratch> rS_s, tS_s;
      if (!same_prec) {
        rS_s.reset(new Scratch(ex.stream, x.X, sp, x.n_parity));
        tS_s.reset(new Scratch(ex.stream, x.X, sp, x.n_parity));
      }
      ColorSpinorField rS = same_prec ? r : rS_s->f;
      ColorSpinorField &xS = xS_s.f, &p = p_s.f, &Ap = Ap_s.f;
      ColorSpinorField tS = same_prec ? tmp : tS_s->f; // sloppy-precision staging

      const double b2 = norm2(b, ex);
      syncs++;
      if (b2 == 0.0) {
        zero(x, ex);
        param.iter = 0;
        param.true_res = 0.0;
        return;
      }
      // r = b - A x
      mat.MdagM(tmp, x);
      copy(r, b, ex);
      double r2 = axpyNorm(-1.0, tmp, r, ex);
      syncs++;
      copy(y, x, ex);
      if (!same_prec) copy(rS, r, ex);
      zero(xS, ex);
      copy(p, rS, ex);
      set_scalar(ex, S_R2, r2);
      const double stop = param.tol * param.tol * b2;
      double rNorm = std::sqrt(r2), r0Norm = rNorm, maxrx = rNorm, maxrr = rNorm;
      int k = 0;
      param.reliable_updates = 0;
      const bool verbose = getenv("B200_CG_VERBOSE") != nullptr;
      if (verbose) fprintf(stderr, "[cg] b2=%g r2=%g stop=%g mixed=%d\n", b2, r2, stop, (int)mixed);

      // fold the sloppy solution into y, recompute the true residual in high precision, repair the recursion
      auto reliable_update = [&]() {
        copy(tmp, xS, ex);
        axpy(1.0, tmp, y, ex);
        zero(xS, ex);
        mat.MdagM(tmp, y);
        copy(r, b, ex);
        const double r2_true = axpyNorm(-1.0, tmp, r, ex);
        syncs++;
        copy(tS, r, ex);             // the true residual in the sloppy precision
        cg_replace_r(p, rS, tS, ex); // p += r_true - rS ; rS = r_true
        set_scalar(ex, S_R2, r2_true);
        param.reliable_updates++;
        return r2_true;
      };
      // convergence / reliable-update decision on a residual norm; returns true if the recursion was restarted
      bool done = false;
      auto decide = [&](double r2_seen) {
        rNorm = std::sqrt(r2_seen);
        if (rNorm > maxrx) maxrx = rNorm;
        if (rNorm > maxrr) maxrr = rNorm;
        const bool converged = r2_seen <= stop;
        const bool update = !same_prec
          && ((rNorm < param.delta * maxrx && r0Norm <= maxrx) || (rNorm < param.delta * r0Norm && r0Norm <= maxrr) || converged);
        if (verbose && (k < 10 || k % 20 == 0 || update)) fprintf(stderr, "[cg] k=%d r2=%g update=%d\n", k, r2_seen, (int)update);
        if (update) {
          r2 = reliable_update();
          rNorm = std::sqrt(r2);
          maxrr = maxrx = r0Norm = rNorm;
          if (r2 <= stop) done = true;
          return true;
        }
        if (converged) done = true;
        return false;
      };

      Pending prev {};
      bool have_prev = false;
      double scratch[S_COUNT];
      while (!done && k < param.maxiter) {
        matSloppy.MdagM(Ap, p);
        Pending pd = cg_dot(p, Ap, ex);
        if (host_ar) { // no NVLink mailboxes: every global sum goes through the host callback (two syncs per iteration)
          const double pAp = await(ex, pd, true, scratch)[S_RAW0];
          syncs++;
          set_scalar(ex, S_PAP, pAp);
          set_scalar(ex, S_ALPHA, r2 / pAp);
        }
        Pending pr = cg_update_r(rS, Ap, ex);
        if (host_ar) {
          const double r2_new = await(ex, pr, true, scratch)[S_RAW0];
          syncs++;
          set_scalar(ex, S_R2_OLD, r2);
          set_scalar(ex, S_R2, r2_new);
          set_scalar(ex, S_BETA, r2_new / r2);
          r2 = r2_new;
        }
        cg_update_xp(xS, p, rS, ex);
        k++;
        if (host_ar) {
          decide(r2);
          continue;
        }
        // the host follows one iteration behind: while the GPU runs iteration k it looks at the residual of k - 1
        bool restarted = false;
        if (have_prev) {
          r2 = await(ex, prev, false, scratch)[S_R2];
          syncs++;
          restarted = decide(r2);
        }
        if (restarted) {
          have_prev = false; // the pending residual belongs to the recursion before the restart
        } else {
          prev = pr;
          have_prev = true;
        }
      }
      // x = y + xS
      if (same_prec) {
        axpy(1.0, xS, y, ex);
      } else {
        copy(tmp, xS, ex);
        axpy(1.0, tmp, y, ex);
      }
      copy(x, y, ex);
      // true residual
      mat.MdagM(tmp, x);
      copy(r, b, ex);
      const double tr2 = axpyNorm(-1.0, tmp, r, ex);
      syncs++;
      cuda_ok(cudaStreamSynchronize(cs(ex.stream)), "sync");
      if (halo_timed_out(ex.comm, ex.stream)) throw Error("a halo wait timed out during the solve: the result is not valid");
      param.iter = k;
      param.true_res = std::sqrt(tr2 / b2);
      param.host_syncs = syncs;
      param.secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const long long nds = mat.DslashApplications() + matSloppy.DslashApplications() - ds0;
      const double fl = (double)(blas::flops() - flops0) + (double)nds * 1320.0 * x.VolumeCB();
      param.gflops = fl / param.secs * 1e-9;
    }

  } // namespace host
} // namespace b200
