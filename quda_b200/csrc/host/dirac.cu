// Implementation of the host-side operator / blas / solver mirror (see dirac.h for the reference citations).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
      if (n_parity != 2) throw Error("parity_view of a single-parity field");
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

    // Field temporaries (reference: getFieldTmp / lib/field_cache.cpp): operators need a scratch field per application;
    // cudaMalloc per call would serialise the stream, so released temporaries are parked in a free list and reused.
    namespace
    {
      struct TmpPool {
        std::vector<ColorSpinorField> free_list;
        ColorSpinorField get(const int *X, int precision, int n_parity)
        {
          for (size_t i = 0; i < free_list.size(); i++) {
            auto &f = free_list[i];
            if (f.precision == precision && f.n_parity == n_parity && f.X[0] == X[0] && f.X[1] == X[1] && f.X[2] == X[2]
                && f.X[3] == X[3]) {
              ColorSpinorField r = f;
              free_list.erase(free_list.begin() + i);
              return r;
            }
          }
          return ColorSpinorField::create(X, precision, n_parity);
        }
        void put(const ColorSpinorField &f) { free_list.push_back(f); }
      };
      TmpPool &pool()
      {
        static TmpPool p;
        return p;
      }
      struct FieldTmp {
        ColorSpinorField f;
        FieldTmp(const ColorSpinorField &like, int n_parity) : f(pool().get(like.X, like.precision, n_parity)) { }
        ~FieldTmp() { pool().put(f); }
        operator ColorSpinorField &() { return f; }
      };
    } // namespace

    // ------------------------------------------------------------------ Apply*
    static void halo_fill(b200_halo &h, const int *comm_override, const CommContext *comm)
    {
      memset(&h, 0, sizeof(h));
      if (!comm) return;
      const int b = comm->seq & 1;
      for (int d = 0; d < 4; d++) {
        h.comm_dim[d] = comm->comm_dim[d] && (!comm_override || comm_override[d]);
        for (int dir = 0; dir < 2; dir++) {
          h.ghost[d][dir] = h.comm_dim[d] ? comm->recv[b][d][dir] : nullptr;
          h.wait_flag[d][dir] = h.comm_dim[d] ? comm->recv_flag[b][d][dir] : nullptr;
        }
      }
      h.seq = comm->seq;
      h.timeout_flag = comm->timeout_flag;
    }

    // ship the faces of `in` (single-parity field holding parity `in_parity`) to the neighbours
    static void exchange_start(const ColorSpinorField &in, int in_parity, bool dagger, const int *comm_override,
                               CommContext *comm, void *stream)
    {
      comm->seq++;
      const int b = comm->seq & 1;
      b200_pack_args a;
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
      a.seq = comm->seq;
      if (comm->pack_stream && comm->pack_stream != stream) {
        // fork: the pack kernel runs on its own stream, concurrently with the interior tiles (joined in apply())
        static cudaEvent_t fork_ev = nullptr;
        if (!fork_ev) cuda_ok(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming), "event");
        cuda_ok(cudaEventRecord(fork_ev, (cudaStream_t)stream), "record");
        cuda_ok(cudaStreamWaitEvent((cudaStream_t)comm->pack_stream, fork_ev, 0), "wait");
        a.stream = comm->pack_stream;
      } else {
        a.stream = stream;
      }
      abi_ok(b200_pack_ghost(&a));
    }

    // `b`, `asymmetric`: twisted mass only; with_x: 1 / 0 force x on / off (the twisted-mass preconditioned operator's
    // xpay flag), -1 = the Wilson convention (x iff a != 0)
    static void apply(int op, ColorSpinorField &out, const ColorSpinorField &in, const GaugeField &U, const CloverField *A,
                      bool inverse_field, double a, const ColorSpinorField &x, int parity, bool dagger,
                      const int *comm_override, CommContext *comm, void *stream, double b = 0.0, bool asymmetric = false,
                      int with_x = -1)
    {
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
      if (part) {
        if (in.n_parity != 1) throw Error("partitioned full-field Dslash: apply per parity (Dirac::M does)");
        exchange_start(in, 1 - parity, dagger, comm_override, comm, stream);
      }
      halo_fill(args.halo, comm_override, part ? comm : nullptr);
      const bool two_streams = part && comm->pack_stream && comm->pack_stream != stream;
      if (two_streams) {
        // side stream (behind the pack kernel): boundary tiles -- they depend only on the halo, not on the interior
        // launch; main stream: the interior tiles.  Both halves write disjoint sites.
        args.kernel = B200_KERNEL_BOUNDARY_TILES;
        args.stream = comm->pack_stream;
        abi_ok(b200_dslash_apply(&args));
        args.kernel = B200_KERNEL_INTERIOR_TILES;
        args.stream = stream;
        abi_ok(b200_dslash_apply(&args));
        // join: `out` is complete, and `in` may be overwritten, only after the side stream has drained
        static cudaEvent_t join_ev = nullptr;
        if (!join_ev) cuda_ok(cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming), "event");
        cuda_ok(cudaEventRecord(join_ev, (cudaStream_t)comm->pack_stream), "record");
        cuda_ok(cudaStreamWaitEvent((cudaStream_t)stream, join_ev, 0), "wait");
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

    // ------------------------------------------------------------------ blas
    namespace blas
    {
      static long long g_flops = 0;
      long long flops() { return g_flops; }

      static double *reduce_buf()
      {
        static double *d = nullptr;
        if (!d) cuda_ok(cudaMalloc(&d, 4 * sizeof(double)), "cudaMalloc(reduce)");
        return d;
      }

      template <typename T> __device__ __forceinline__ double blk_sum(double v, double *out)
      {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        __shared__ double s[32];
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) s[w] = v;
        __syncthreads();
        if (w == 0) {
          v = l < (blockDim.x >> 5) ? s[l] : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
          if (l == 0) atomicAdd(out, v);
        }
        return v;
      }

      // ---- NVLink mailbox all-reduce of up to 4 doubles (b200_comm::reduce_peer).  One warp: lane r < n_ranks pushes this
      // rank's partial sums into its slot of rank r's mailbox and raises the slot's sequence number, then waits for
      // rank r's contribution in the local mailbox; lane 0 adds the contributions in rank order (identical on every
      // rank).  Slots are double buffered by the parity of `seq`: a rank can only be one reduction ahead of the
      // slowest one, because finishing reduction k needs everybody's contribution to k.
      struct ReduceSlot {
        double v[4];
        unsigned seq;
        unsigned pad[7];
      };
      static_assert(sizeof(ReduceSlot) == B200_REDUCE_SLOT_BYTES, "mailbox slot layout");
      struct ReducePeers {
        ReduceSlot *box[B200_MAX_RANKS];
      };

      __global__ void mailbox_allreduce_kernel(double *__restrict__ val, int n, ReducePeers peers, int rank, int n_ranks,
                                               unsigned seq, int *timeout_flag)
      {
        __shared__ double part[B200_MAX_RANKS][4];
        const int t = threadIdx.x;
        const int b = seq & 1;
        if (t < n_ranks) {
          ReduceSlot *dst = peers.box[t] + b * B200_MAX_RANKS + rank;
          for (int i = 0; i < n; i++) dst->v[i] = val[i];
          __threadfence_system();
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&dst->seq), "r"(seq) : "memory");
          const ReduceSlot *src = peers.box[rank] + b * B200_MAX_RANKS + t;
          const long long t0 = clock64();
          for (;;) {
            unsigned got;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(&src->seq) : "memory");
            if ((int)(got - seq) >= 0) break;
            if (clock64() - t0 > 4000000000LL) { // ~2 s: a lost peer must never hang the GPU
              if (timeout_flag) *timeout_flag = 1;
              break;
            }
            __nanosleep(50);
          }
          for (int i = 0; i < n; i++) part[t][i] = *reinterpret_cast<const volatile double *>(&src->v[i]);
        }
        __syncthreads();
        if (t == 0) {
          for (int i = 0; i < n; i++) {
            double acc = 0;
            for (int r = 0; r < n_ranks; r++) acc += part[r][i];
            val[i] = acc;
          }
        }
      }

      // sum `n` device doubles over all ranks in place (stream-ordered); false if this CommContext has no mailboxes
      static bool device_allreduce(double *val, int n, CommContext *comm)
      {
        if (!comm || comm->n_ranks < 2) return false;
        if (n > 4 || comm->n_ranks > B200_MAX_RANKS) throw Error("mailbox all-reduce: too many values / ranks");
        ReducePeers peers;
        for (int r = 0; r < B200_MAX_RANKS; r++) peers.box[r] = reinterpret_cast<ReduceSlot *>(comm->reduce_peer[r]);
        comm->reduce_seq++;
        mailbox_allreduce_kernel<<<1, 32>>>(val, n, peers, comm->rank, comm->n_ranks, comm->reduce_seq, comm->timeout_flag);
        cuda_ok(cudaGetLastError(), "all-reduce launch");
        return true;
      }

      // y = a x + b y (+ optional z update), optional reduction of |y|^2 or <x,y>; one template keeps it compact
      enum { R_NONE = 0, R_NORM_Y = 1, R_DOT_XY = 2 };
      template <typename Tx, typename Ty, int R>
      __global__ void axpby_kernel(double a, const Tx *__restrict__ x, double b, Ty *__restrict__ y, size_t n, double *red,
                                   bool write)
      {
        double acc = 0;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
          const double xv = x[i], yv = y[i];
          const double r = a * xv + b * yv;
          if (write) y[i] = (Ty)r;
          if (R == R_NORM_Y) acc += (write ? (double)(Ty)r * (double)(Ty)r : yv * yv);
          if (R == R_DOT_XY) acc += xv * yv;
        }
        if (R != R_NONE) blk_sum<double>(acc, red);
      }

      template <typename T>
      __global__ void axpyZpbx_kernel(double a, T *__restrict__ p, T *__restrict__ x, const T *__restrict__ r, double b, size_t n)
      {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
          const double pv = p[i];
          x[i] = (T)((double)x[i] + a * pv);
          p[i] = (T)((double)r[i] + b * pv);
        }
      }

      static void check_pair(const ColorSpinorField &x, const ColorSpinorField &y)
      {
        if (x.Length() != y.Length()) throw Error("blas: field length mismatch");
        if (x.precision == B200_HALF || y.precision == B200_HALF) throw Error("blas: block-float half fields not supported");
      }

      template <int R>
      static double run(double a, const ColorSpinorField &x, double b, ColorSpinorField &y, bool write, CommContext *comm)
      {
        check_pair(x, y);
        if (x.precision != y.precision) throw Error("blas: mixed-precision operands (use blas::copy to convert)");
        const size_t n = x.Length();
        double *red = reduce_buf();
        if (R != R_NONE) cuda_ok(cudaMemsetAsync(red, 0, sizeof(double)), "memset");
        const int threads = 256, blocks = 148 * 8;
        if (x.precision == 8 && y.precision == 8)
          axpby_kernel<double, double, R><<<blocks, threads>>>(a, (const double *)x.v, b, (double *)y.v, n, red, write);
        else if (x.precision == 4 && y.precision == 4)
          axpby_kernel<float, float, R><<<blocks, threads>>>(a, (const float *)x.v, b, (float *)y.v, n, red, write);
        else if (x.precision == 8 && y.precision == 4)
          axpby_kernel<double, float, R><<<blocks, threads>>>(a, (const double *)x.v, b, (float *)y.v, n, red, write);
        else
          axpby_kernel<float, double, R><<<blocks, threads>>>(a, (const float *)x.v, b, (double *)y.v, n, red, write);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 3 * (long long)n;
        if (R == R_NONE) return 0.0;
        double h = 0;
        const bool on_device = device_allreduce(red, 1, comm); // NVLink mailboxes if wired, else the host callback
        cuda_ok(cudaMemcpy(&h, red, sizeof(double), cudaMemcpyDeviceToHost), "memcpy(reduce)");
        if (!on_device && comm && comm->allreduce_sum) comm->allreduce_sum(&h, 1, comm->user);
        return h;
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

      void copy(ColorSpinorField &dst, const ColorSpinorField &src)
      {
        check_pair(src, dst);
        if (dst.n_parity != src.n_parity) throw Error("copy: site subsets differ");
        if (src.precision == dst.precision) {
          cuda_ok(cudaMemcpyAsync(dst.v, src.v, src.Bytes(), cudaMemcpyDeviceToDevice), "copy");
          return;
        }
        const int vcb = src.VolumeCB();
        dim3 grid((vcb + 127) / 128, src.n_parity);
        const size_t se = (size_t)24 * vcb, de = (size_t)24 * vcb;
        if (src.precision == 8)
          convert_kernel<double, 2, float, 4><<<grid, 128>>>((const double *)src.v, (float *)dst.v, vcb, se, de);
        else
          convert_kernel<float, 4, double, 2><<<grid, 128>>>((const float *)src.v, (double *)dst.v, vcb, se, de);
        cuda_ok(cudaGetLastError(), "convert launch");
      }
      void zero(ColorSpinorField &x) { cuda_ok(cudaMemsetAsync(x.v, 0, x.Bytes()), "memset"); }
      void ax(double a, ColorSpinorField &x) { run<R_NONE>(0.0, x, a, x, true, nullptr); }
      void axpy(double a, const ColorSpinorField &x, ColorSpinorField &y) { run<R_NONE>(a, x, 1.0, y, true, nullptr); }
      void xpay(const ColorSpinorField &x, double a, ColorSpinorField &y) { run<R_NONE>(1.0, x, a, y, true, nullptr); }
      void axpby(double a, const ColorSpinorField &x, double b, ColorSpinorField &y) { run<R_NONE>(a, x, b, y, true, nullptr); }
      double norm2(const ColorSpinorField &x, CommContext *comm)
      {
        return run<R_NORM_Y>(0.0, x, 1.0, const_cast<ColorSpinorField &>(x), false, comm);
      }
      double reDotProduct(const ColorSpinorField &x, const ColorSpinorField &y, CommContext *comm)
      {
        return run<R_DOT_XY>(0.0, x, 1.0, const_cast<ColorSpinorField &>(y), false, comm);
      }
      double axpyNorm(double a, const ColorSpinorField &x, ColorSpinorField &y, CommContext *comm)
      {
        return run<R_NORM_Y>(a, x, 1.0, y, true, comm);
      }
      double xmyNorm(const ColorSpinorField &x, ColorSpinorField &y, CommContext *comm)
      {
        return run<R_NORM_Y>(1.0, x, -1.0, y, true, comm);
      }
      void axpyZpbx(double a, ColorSpinorField &p, ColorSpinorField &x, const ColorSpinorField &r, double b)
      {
        check_pair(p, x);
        check_pair(p, r);
        if (p.precision != x.precision || p.precision != r.precision) throw Error("axpyZpbx: mixed precision");
        const size_t n = p.Length();
        if (p.precision == 8)
          axpyZpbx_kernel<double><<<148 * 8, 256>>>(a, (double *)p.v, (double *)x.v, (const double *)r.v, b, n);
        else
          axpyZpbx_kernel<float><<<148 * 8, 256>>>(a, (float *)p.v, (float *)x.v, (const float *)r.v, b, n);
        cuda_ok(cudaGetLastError(), "blas launch");
        g_flops += 4 * (long long)n;
      }
    } // namespace blas

    // ------------------------------------------------------------------ Dirac
    Dirac::Dirac(const DiracParam &p) :
      gauge(p.gauge), kappa(p.kappa), matpcType(p.matpcType), dagger(p.dagger), comm(p.comm), stream(p.stream)
    {
      if (!gauge) throw Error("Dirac: gauge field missing");
      for (int d = 0; d < 4; d++) commDim[d] = p.commDim[d];
      symmetric = (matpcType == QUDA_MATPC_EVEN_EVEN || matpcType == QUDA_MATPC_ODD_ODD);
      this_parity = (matpcType == QUDA_MATPC_EVEN_EVEN || matpcType == QUDA_MATPC_EVEN_EVEN_ASYMMETRIC) ? 0 : 1;
      other_parity = 1 - this_parity;
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

    Dirac *Dirac::create(const std::string &type, const DiracParam &p)
    {
      if (type == "wilson") return new DiracWilson(p);
      if (type == "wilsonpc") return new DiracWilsonPC(p);
      if (type == "clover") return new DiracClover(p);
      if (type == "cloverpc") return new DiracCloverPC(p);
      if (type == "twistedmass") return new DiracTwistedMass(p);
      if (type == "twistedmasspc") return new DiracTwistedMassPC(p);
      throw Error("Dirac::create: unsupported operator type '" + type + "'");
    }

    static void check_parity_spinor(const ColorSpinorField &a, const ColorSpinorField &b)
    {
      if (a.n_parity != 1 || b.n_parity != 1) throw Error("ColorSpinorFields are not single parity");
      if (a.v == b.v) throw Error("Aliasing pointers");
    }
    static void check_full_spinor(const ColorSpinorField &a, const ColorSpinorField &b)
    {
      if (a.n_parity != 2 || b.n_parity != 2) throw Error("ColorSpinorFields are not full fields");
    }

    // --- Wilson (lib/dirac_wilson.cpp:21-104)
    void DiracWilson::Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const
    {
      check_parity_spinor(in, out);
      ApplyWilson(out, in, *gauge, 0.0, in, parity, dagger, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracWilson::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                                 double k) const
    {
      check_parity_spinor(in, out);
      ApplyWilson(out, in, *gauge, k, x, parity, dagger, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracWilson::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      check_full_spinor(out, in);
      if (comm && comm->partitioned()) { // halo exchange works per parity
        auto oe = out.Even(), oo = out.Odd();
        DiracWilson::DslashXpay(oe, in.Odd(), 0, in.Even(), -kappa);
        DiracWilson::DslashXpay(oo, in.Even(), 1, in.Odd(), -kappa);
      } else {
        ApplyWilson(out, in, *gauge, -kappa, in, QUDA_INVALID_PARITY, dagger, commDim, comm, stream);
        dslash_applications += 2;
      }
    }
    void DiracWilson::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      FieldTmp tmp(in, in.n_parity);
      M(tmp, in);
      Mdag(out, tmp);
    }
    void DiracWilson::prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                              QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION)
        throw Error("Preconditioned solution requires a preconditioned solve_type");
      src = b;
      sol = x;
    }
    void DiracWilson::reconstruct(ColorSpinorField &, const ColorSpinorField &, QudaSolutionType) const { }

    // --- WilsonPC (lib/dirac_wilson.cpp:106-163)
    void DiracWilsonPC::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      const double kappa2 = -kappa * kappa;
      FieldTmp tmp(in, 1);
      if (!symmetric) throw Error("MatPCType not valid for DiracWilsonPC");
      Dslash(tmp, in, other_parity);
      DslashXpay(out, tmp, this_parity, in, kappa2);
    }
    void DiracWilsonPC::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      FieldTmp tmp(in, 1);
      M(tmp, in);
      Mdag(out, tmp);
    }
    void DiracWilsonPC::prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                                QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) {
        src = b;
        sol = x;
        return;
      }
      // src = b_e + k D_eo b_o (stored in x_o), solution in x_e
      auto xo = x.parity_view(other_parity);
      DslashXpay(xo, b.parity_view(other_parity), this_parity, b.parity_view(this_parity), kappa);
      src = xo;
      sol = x.parity_view(this_parity);
    }
    void DiracWilsonPC::reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) return;
      check_full_spinor(x, b);
      auto xo = x.parity_view(other_parity);
      DslashXpay(xo, x.parity_view(this_parity), other_parity, b.parity_view(other_parity), kappa);
    }

    // --- Clover (lib/dirac_clover.cpp:36-100)
    DiracClover::DiracClover(const DiracParam &p) : DiracWilson(p), clover(p.clover)
    {
      if (!clover) throw Error("DiracClover: clover field missing");
    }
    void DiracClover::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                                 double k) const
    {
      check_parity_spinor(in, out);
      ApplyWilsonClover(out, in, *gauge, *clover, k, x, parity, dagger, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracClover::Clover(ColorSpinorField &out, const ColorSpinorField &in, int parity) const
    {
      ApplyClover(out, in, *clover, false, parity, stream);
    }
    void DiracClover::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      check_full_spinor(out, in);
      if (comm && comm->partitioned()) {
        auto oe = out.Even(), oo = out.Odd();
        DiracClover::DslashXpay(oe, in.Odd(), 0, in.Even(), -kappa);
        DiracClover::DslashXpay(oo, in.Even(), 1, in.Odd(), -kappa);
      } else {
        ApplyWilsonClover(out, in, *gauge, *clover, -kappa, in, QUDA_INVALID_PARITY, dagger, commDim, comm, stream);
        dslash_applications += 2;
      }
    }
    void DiracClover::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      check_full_spinor(out, in);
      FieldTmp tmp(in, 2);
      M(tmp, in);
      Mdag(out, tmp);
    }

    // --- CloverPC (lib/dirac_clover.cpp:118-258)
    DiracCloverPC::DiracCloverPC(const DiracParam &p) : DiracClover(p)
    {
      if (!clover->has_inverse() && !clover->c.dynamic_inverse) throw Error("Clover inverse required for DiracCloverPC");
    }
    void DiracCloverPC::CloverInv(ColorSpinorField &out, const ColorSpinorField &in, int parity) const
    {
      ApplyClover(out, in, *clover, true, parity, stream);
    }
    void DiracCloverPC::Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const
    {
      check_parity_spinor(in, out);
      ApplyWilsonCloverPreconditioned(out, in, *gauge, *clover, 0.0, in, parity, dagger, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracCloverPC::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                                   double k) const
    {
      check_parity_spinor(in, out);
      ApplyWilsonCloverPreconditioned(out, in, *gauge, *clover, k, x, parity, dagger, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracCloverPC::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      const double kappa2 = -kappa * kappa;
      FieldTmp tmp(in, 1);
      if (!symmetric) {
        Dslash(tmp, in, other_parity);                                   // A^-1 D
        DiracClover::DslashXpay(out, tmp, this_parity, in, kappa2);      // A x - k^2 D
      } else if (!dagger) {
        Dslash(tmp, in, other_parity);
        DslashXpay(out, tmp, this_parity, in, kappa2);                   // x - k^2 A^-1 D (A^-1 D)
      } else {
        CloverInv(out, in, this_parity);                                 // 1 - D^+ A^-1 D^+ A^-1
        Dslash(tmp, out, other_parity);
        DiracWilson::DslashXpay(out, tmp, this_parity, in, kappa2);
      }
    }
    void DiracCloverPC::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      FieldTmp tmp(in, 1);
      M(tmp, in);
      Mdag(out, tmp);
    }
    void DiracCloverPC::prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                                QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) {
        src = b;
        sol = x;
        return;
      }
      src = x.parity_view(other_parity);
      sol = x.parity_view(this_parity);
      FieldTmp tmp(b, 1);
      if (symmetric) { // src = A_ee^-1 (b_e + k D_eo A_oo^-1 b_o)
        CloverInv(src, b.parity_view(other_parity), other_parity);
        DiracWilson::DslashXpay(tmp, src, this_parity, b.parity_view(this_parity), kappa);
        CloverInv(src, tmp, this_parity);
      } else { // src = b_e + k D_eo A_oo^-1 b_o
        CloverInv(tmp, b.parity_view(other_parity), other_parity);
        DiracWilson::DslashXpay(src, tmp, this_parity, b.parity_view(this_parity), kappa);
      }
    }
    void DiracCloverPC::reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) return;
      check_full_spinor(x, b);
      FieldTmp tmp(b, 1);
      // x_o = A_oo^-1 (b_o + k D_oe x_e)
      DiracWilson::DslashXpay(tmp, x.parity_view(this_parity), other_parity, b.parity_view(other_parity), kappa);
      auto xo = x.parity_view(other_parity);
      CloverInv(xo, tmp, other_parity);
    }

    // --- twisted mass, singlet flavour (lib/dirac_twisted_mass.cpp:9-317)
    DiracTwistedMass::DiracTwistedMass(const DiracParam &p) : DiracWilson(p), mu(p.mu) { }
    void DiracTwistedMass::Twist(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      ApplyTwistGamma(out, in, kappa, mu, dagger, false, stream);
    }
    void DiracTwistedMass::Dslash(ColorSpinorField &, const ColorSpinorField &, int) const
    {
      // the reference routes this to ApplyTwistedMass with a = 0, which it does not instantiate (:47-59)
      throw Error("DiracTwistedMass::Dslash: twisted-mass operator only defined for xpay=true");
    }
    void DiracTwistedMass::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                                      double k) const
    {
      check_parity_spinor(in, out);
      ApplyTwistedMass(out, in, *gauge, k, 2 * mu * kappa, x, parity, dagger, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracTwistedMass::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      check_full_spinor(out, in);
      // -kappa D in + (1 + i 2 mu kappa gamma5) in, one parity at a time (the halo exchange works per parity)
      auto oe = out.Even(), oo = out.Odd();
      DiracTwistedMass::DslashXpay(oe, in.Odd(), 0, in.Even(), -kappa);
      DiracTwistedMass::DslashXpay(oo, in.Even(), 1, in.Odd(), -kappa);
    }
    void DiracTwistedMass::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      FieldTmp tmp(in, in.n_parity);
      M(tmp, in);
      Mdag(out, tmp);
    }

    void DiracTwistedMassPC::TwistInv(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      ApplyTwistGamma(out, in, kappa, mu, dagger, true, stream);
    }
    void DiracTwistedMassPC::Dslash(ColorSpinorField &out, const ColorSpinorField &in, int parity) const
    {
      check_parity_spinor(in, out);
      const double a = -2.0 * kappa * mu; // inverse twist
      const double b = 1.0 / (1.0 + a * a);
      const bool asymmetric = !symmetric && dagger;
      ApplyTwistedMassPreconditioned(out, in, *gauge, b, a, false, in, parity, dagger, asymmetric, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracTwistedMassPC::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, int parity, const ColorSpinorField &x,
                                        double k) const
    {
      check_parity_spinor(in, out);
      const double a = -2.0 * kappa * mu;
      const double b = k / (1.0 + a * a);
      const bool asymmetric = !symmetric && dagger;
      ApplyTwistedMassPreconditioned(out, in, *gauge, b, a, true, x, parity, dagger, asymmetric, commDim, comm, stream);
      dslash_applications++;
    }
    void DiracTwistedMassPC::M(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      const double kappa2 = -kappa * kappa;
      FieldTmp tmp(in, 1);
      Dslash(tmp, in, other_parity);
      if (symmetric)
        DslashXpay(out, tmp, this_parity, in, kappa2);
      else
        DiracTwistedMass::DslashXpay(out, tmp, this_parity, in, kappa2);
    }
    void DiracTwistedMassPC::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const
    {
      FieldTmp tmp(in, 1); // extra temporary because of the symmetric dagger operator
      M(tmp, in);
      Mdag(out, tmp);
    }
    void DiracTwistedMassPC::prepare(ColorSpinorField &sol, ColorSpinorField &src, ColorSpinorField &x, const ColorSpinorField &b,
                                     QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) {
        src = b;
        sol = x;
        return;
      }
      src = x.parity_view(other_parity);
      sol = x.parity_view(this_parity);
      FieldTmp tmp(b, 1);
      if (symmetric) { // src = A_ee^-1 (b_e + k D_eo A_oo^-1 b_o)
        TwistInv(src, b.parity_view(other_parity));
        DiracWilson::DslashXpay(tmp, src, this_parity, b.parity_view(this_parity), kappa);
        TwistInv(src, tmp);
      } else { // src = b_e + k D_eo A_oo^-1 b_o
        TwistInv(tmp, b.parity_view(other_parity));
        DiracWilson::DslashXpay(src, tmp, this_parity, b.parity_view(this_parity), kappa);
      }
    }
    void DiracTwistedMassPC::reconstruct(ColorSpinorField &x, const ColorSpinorField &b, QudaSolutionType st) const
    {
      if (st == QUDA_MATPC_SOLUTION || st == QUDA_MATPCDAG_MATPC_SOLUTION) return;
      check_full_spinor(x, b);
      FieldTmp tmp(b, 1);
      // x_o = A_oo^-1 (b_o + k D_oe x_e)
      DiracWilson::DslashXpay(tmp, x.parity_view(this_parity), other_parity, b.parity_view(other_parity), kappa);
      auto xo = x.parity_view(other_parity);
      TwistInv(xo, tmp);
    }

    // ------------------------------------------------------------------ CG (normal equations) with reliable updates
    void invertCG(const Dirac &mat, const Dirac &matSloppy, ColorSpinorField &x, const ColorSpinorField &b, SolverParam &param)
    {
      using namespace blas;
      CommContext *comm = mat.Comm();
      const auto t0 = std::chrono::steady_clock::now();
      const long long flops0 = blas::flops();
      const long long ds0 = mat.DslashApplications() + matSloppy.DslashApplications();
      const bool mixed = (&mat != &matSloppy);
      const int sp = mixed ? 4 : x.precision; // sloppy precision
      if (mixed && x.precision != 8) throw Error("mixed-precision CG expects a double-precision solution field");

      auto r = ColorSpinorField::create(x.X, x.precision, x.n_parity);  // high-precision residual
      auto y = ColorSpinorField::create(x.X, x.precision, x.n_parity);  // high-precision accumulated solution
      auto tmp = ColorSpinorField::create(x.X, x.precision, x.n_parity);
      auto rS = mixed ? ColorSpinorField::create(x.X, sp, x.n_parity) : r;
      auto xS = ColorSpinorField::create(x.X, sp, x.n_parity);
      auto p = ColorSpinorField::create(x.X, sp, x.n_parity);
      auto Ap = ColorSpinorField::create(x.X, sp, x.n_parity);

      const double b2 = norm2(b, comm);
      if (b2 == 0.0) {
        zero(x);
        param.iter = 0;
        param.true_res = 0.0;
        return;
      }
      // r = b - A x
      mat.MdagM(tmp, x);
      copy(r, b);
      axpy(-1.0, tmp, r);
      double r2 = norm2(r, comm);
      copy(y, x);
      if (mixed) copy(rS, r);
      zero(xS);
      copy(p, rS);
      const double stop = param.tol * param.tol * b2;
      double rNorm = std::sqrt(r2), r0Norm = rNorm, maxrx = rNorm, maxrr = rNorm;
      int k = 0;
      param.reliable_updates = 0;
      const bool verbose = getenv("B200_CG_VERBOSE") != nullptr;
      if (verbose) fprintf(stderr, "[cg] b2=%g r2=%g stop=%g mixed=%d\n", b2, r2, stop, (int)mixed);
      while (r2 > stop && k < param.maxiter) {
        matSloppy.MdagM(Ap, p);
        const double pAp = reDotProduct(p, Ap, comm);
        const double alpha = r2 / pAp;
        const double r2_old = r2;
        r2 = axpyNorm(-alpha, Ap, rS, comm);
        rNorm = std::sqrt(r2);
        if (rNorm > maxrx) maxrx = rNorm;
        if (rNorm > maxrr) maxrr = rNorm;
        const bool update = mixed && ((rNorm < param.delta * maxrx && r0Norm <= maxrx) || (rNorm < param.delta * r0Norm && r0Norm <= maxrr) || r2 <= stop);
        if (!update) {
          const double beta = r2 / r2_old;
          axpyZpbx(alpha, p, xS, rS, beta); // xS += alpha p ; p = rS + beta p
        } else {
          axpy(alpha, p, xS);
          // reliable update: fold the sloppy solution into y, recompute the true residual in high precision
          copy(tmp, xS);
          axpy(1.0, tmp, y);
          mat.MdagM(tmp, y);
          copy(r, b);
          r2 = axpyNorm(-1.0, tmp, r, comm);
          copy(rS, r);
          zero(xS);
          // keep the search direction conjugate as far as precision allows: p = r + beta p
          const double beta = r2 / r2_old;
          xpay(rS, beta, p);
          rNorm = std::sqrt(r2);
          maxrr = maxrx = r0Norm = rNorm;
          param.reliable_updates++;
        }
        k++;
        if (verbose && (k < 10 || k % 20 == 0 || update))
          fprintf(stderr, "[cg] k=%d r2=%g pAp=%g alpha=%g update=%d\n", k, r2, pAp, alpha, (int)update);
      }
      // x = y + xS
      copy(tmp, xS);
      axpy(1.0, tmp, y);
      copy(x, y);
      // true residual
      mat.MdagM(tmp, x);
      copy(r, b);
      const double tr2 = axpyNorm(-1.0, tmp, r, comm);
      cuda_ok(cudaDeviceSynchronize(), "sync");
      param.iter = k;
      param.true_res = std::sqrt(tr2 / b2);
      param.secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const long long nds = mat.DslashApplications() + matSloppy.DslashApplications() - ds0;
      const double fl = (double)(blas::flops() - flops0) + (double)nds * 1320.0 * x.VolumeCB();
      param.gflops = fl / param.secs * 1e-9;
    }

  } // namespace host
} // namespace b200
