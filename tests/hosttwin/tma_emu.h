// TEST-ONLY emulation of the TMA-staged marching kernel (quda_b200/csrc/tma_kernel.cuh) on the CPU.
//
// It runs the product's plan, tensor-map descriptions, box coordinates, per-thread shared-memory offsets, producer
// program and site arithmetic (quda_b200/csrc/tma.h) against an emulated shared memory:
//   * box loads are interpreted from the TmaDesc exactly as cuTensorMapEncodeTiled / cp.async.bulk.tensor define them
//     (dense box in dimension order, out-of-bounds elements zero-filled);
//   * producer and consumers run as two programs coupled only through ring-slot ownership (the mbarrier full / empty
//     protocol): the producer runs as far ahead as the slots allow, a slot may only be overwritten after the consumers
//     released it, released slots are poisoned, and a state in which neither side can advance is reported as a deadlock.
// What it cannot check: PTX syntax and the phase-parity arithmetic of the real mbarriers (GPU tests do).
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

#include "../../quda_b200/csrc/tma.h"

namespace b200
{

  inline void tma_emu_box_load(unsigned char *dst, const TmaDesc &d, const int *c)
  {
    size_t o = 0;
    const unsigned char *base = static_cast<const unsigned char *>(d.base);
    for (unsigned i4 = 0; i4 < d.box[4]; i4++)
      for (unsigned i3 = 0; i3 < d.box[3]; i3++)
        for (unsigned i2 = 0; i2 < d.box[2]; i2++)
          for (unsigned i1 = 0; i1 < d.box[1]; i1++)
            for (unsigned i0 = 0; i0 < d.box[0]; i0++, o += 4) {
              const long long x[5] = {c[0] + (long long)i0, c[1] + (long long)i1, c[2] + (long long)i2, c[3] + (long long)i3, c[4] + (long long)i4};
              bool in = true;
              unsigned long long off = 0;
              for (int k = 0; k < 5; k++) {
                if (x[k] < 0 || (unsigned long long)x[k] >= d.dim[k]) in = false;
                off += (unsigned long long)x[k] * d.stride[k];
              }
              if (in)
                memcpy(dst + o, base + off, 4);
              else
                memset(dst + o, 0, 4);
            }
  }

  struct TmaEmuOp {
    int ring; // 0 centre, 1 halo, 2 link
    long long n;
    TmaItem it;
    int arg; // slice (centre) or d (link)
  };

  struct TmaEmuRecorder {
    std::vector<TmaEmuOp> ops;
    void center(int n, const TmaItem &it, int slice) { ops.push_back({0, n, it, slice}); }
    void halo(int n, const TmaItem &it) { ops.push_back({1, n, it, 0}); }
    void link(long long ln, const TmaItem &it, int d) { ops.push_back({2, ln, it, d}); }
  };

  // one CTA; returns 0 or an error code (message via set_error)
  template <class P, int recon, bool dagger, bool xpay, OpType op>
  int tma_emu_cta(const DslashArgs<P, recon> &arg, const TmaPlan &plan, const TmaDesc (*descs)[TM_COUNT], int cta, int n_cta, long &visited)
  {
    using real = typename P::real;
    int w0, w1;
    tma_work_range(w0, w1, plan, cta, n_cta);
    if (w0 >= w1) return 0;
    const int NL = plan.n_link_slots;
    std::vector<unsigned char> smem(plan.smem_bytes, 0xA5);
    std::vector<long long> own_c(kTmaCenterSlots, -1), own_h(kTmaHaloSlots, -1), own_l(NL, -1);
    TmaEmuRecorder rec;
    tma_producer_program(plan, w0, w1, rec);
    size_t pi = 0;

    auto poison = [&](int off, int bytes) { memset(smem.data() + off, 0xFF, bytes); }; // NaN pattern for every real type
    auto desc_of = [&](const TmaItem &it, int id) -> const TmaDesc & { return descs[plan.n_parity == 2 ? it.par : 0][id]; };
    // producer: run as far as the slot ownership allows
    auto produce = [&]() -> int {
      int done = 0;
      while (pi < rec.ops.size()) {
        const TmaEmuOp &o = rec.ops[pi];
        if (o.ring == 0) {
          const int s = (int)(o.n & (kTmaCenterSlots - 1));
          if (own_c[s] >= 0) break;
          TmaBox b;
          tma_center_box(b, o.it, o.arg);
          const TmaDesc &d = desc_of(o.it, b.map);
          if (tma_box_bytes(d) != plan.NC * plan.srow) return set_error(B200_ERR_INVALID, "centre box bytes != expect_tx");
          tma_emu_box_load(smem.data() + plan.off_center + s * plan.center_bytes + b.dst, d, b.c);
          own_c[s] = o.n;
        } else if (o.ring == 1) {
          const int s = (int)(o.n & (kTmaHaloSlots - 1));
          if (own_h[s] >= 0) break;
          TmaBox b[4];
          tma_halo_boxes(b, plan, o.it);
          int bytes = 0;
          for (int k = 0; k < 4; k++) {
            const TmaDesc &d = desc_of(o.it, b[k].map);
            bytes += tma_box_bytes(d);
            tma_emu_box_load(smem.data() + plan.off_halo + s * plan.halo_bytes + b[k].dst, d, b[k].c);
          }
          if (bytes != plan.NH * plan.srow) return set_error(B200_ERR_INVALID, "halo box bytes != expect_tx");
          own_h[s] = o.n;
        } else {
          const int s = (int)(o.n % NL);
          if (own_l[s] >= 0) break;
          TmaBox b[3];
          const int nb = tma_link_boxes(b, plan, o.it, o.arg);
          int bytes = 0;
          for (int k = 0; k < nb; k++) {
            const TmaDesc &d = desc_of(o.it, b[k].map);
            if (!d.valid) return set_error(B200_ERR_INVALID, "link box uses an invalid map");
            bytes += tma_box_bytes(d);
            tma_emu_box_load(smem.data() + plan.off_link + s * plan.link_bytes + b[k].dst, d, b[k].c);
          }
          if (bytes != 2 * plan.NC * plan.grow) return set_error(B200_ERR_INVALID, "link box bytes != expect_tx");
          own_l[s] = o.n;
        }
        pi++;
        done++;
      }
      return done;
    };
    // consumer-side wait: the load must be resident; otherwise let the producer run; if it cannot, that is a deadlock
    auto wait = [&](std::vector<long long> &own, int slot, long long n, const char *what) -> int {
      if (own[slot] == n) return 0;
      const int r = produce();
      if (r < 0) return r;
      if (own[slot] != n)
        return set_error(B200_ERR_INVALID, "TMA pipeline deadlock: CTA %d waits for %s load %lld (slot holds %lld, producer at op %zu/%zu)", cta,
                         what, n, own[slot], pi, rec.ops.size());
      return 0;
    };

    const int nthr = plan.n_cwarps * 32;
    std::vector<TmaThread> th(nthr);
    for (int t = 0; t < nthr; t++) tma_thread_init(th[t], plan, t);
    std::vector<real> acc((size_t)nthr * 24);
    int cn = 0, hn = 0;
    long long ln = 0;
    {
      const int r = produce(); // the producer starts before anybody waits
      if (r < 0) return r;
    }
    for (int w = w0; w < w1; w++) {
      TmaItem it;
      tma_item(it, plan, w);
      const bool first = (w == w0) || (it.t == 0);
      const bool last = (w == w1 - 1) || (it.t == plan.T - 1);
      if (first) {
        if (int e = wait(own_c, cn & 3, cn, "centre")) return e;
        if (int e = wait(own_c, (cn + 1) & 3, cn + 1, "centre")) return e;
      }
      if (int e = wait(own_c, (cn + 2) & 3, cn + 2, "centre")) return e;
      if (int e = wait(own_h, hn & 1, hn, "halo")) return e;
      // slices t-1 and t must still be the loads this item expects (they were waited for by earlier items)
      if (own_c[cn & 3] != cn || own_c[(cn + 1) & 3] != cn + 1)
        return set_error(B200_ERR_INVALID, "TMA pipeline: centre slot overwritten while live (CTA %d item %d)", cta, w);
      TmaBases b;
      b.cm = smem.data() + plan.off_center + (cn & 3) * plan.center_bytes;
      b.c0 = smem.data() + plan.off_center + ((cn + 1) & 3) * plan.center_bytes;
      b.cp = smem.data() + plan.off_center + ((cn + 2) & 3) * plan.center_bytes;
      b.halo = smem.data() + plan.off_halo + (hn & 1) * plan.halo_bytes;
      for (size_t i = 0; i < acc.size(); i++) acc[i] = 0;
      auto dim = [&](auto D) -> int {
        constexpr int d = decltype(D)::value;
        const int slot = (int)(ln % NL);
        if (int e = wait(own_l, slot, ln, "link")) return e;
        const unsigned char *stage = smem.data() + plan.off_link + slot * plan.link_bytes;
        for (int t = 0; t < nthr; t++) {
          if (!th[t].active) continue;
          int x_cb, x_cb_tm, rpar;
          tma_site_index(x_cb, x_cb_tm, rpar, plan, th[t], it);
          tma_hop_pair<P, recon, dagger, d>(&acc[(size_t)t * 24], arg.U, plan, th[t], b, stage, rpar, x_cb, x_cb_tm);
        }
        own_l[slot] = -1; // released
        poison(plan.off_link + slot * plan.link_bytes, plan.link_bytes);
        ln++;
        const int r = produce(); // aggressive producer: refill as soon as a slot is free
        return r < 0 ? r : 0;
      };
      if (int e = dim(std::integral_constant<int, 0> {})) return e;
      if (int e = dim(std::integral_constant<int, 1> {})) return e;
      if (int e = dim(std::integral_constant<int, 2> {})) return e;
      if (int e = dim(std::integral_constant<int, 3> {})) return e;
      auto release_c = [&](int n) {
        own_c[n & 3] = -1;
        poison(plan.off_center + (n & 3) * plan.center_bytes, plan.center_bytes);
      };
      release_c(cn);
      own_h[hn & 1] = -1;
      poison(plan.off_halo + (hn & 1) * plan.halo_bytes, plan.halo_bytes);
      if (last) {
        release_c(cn + 1);
        release_c(cn + 2);
      }
      {
        const int r = produce();
        if (r < 0) return r;
      }
      for (int t = 0; t < nthr; t++) {
        if (!th[t].active) continue;
        int x_cb, x_cb_tm, rpar;
        tma_site_index(x_cb, x_cb_tm, rpar, plan, th[t], it);
        tma_epilogue<P, recon, dagger, xpay, op>(&acc[(size_t)t * 24], arg, x_cb, it.par);
        visited++;
      }
      cn += last ? 3 : 1;
      hn++;
    }
    if (pi != rec.ops.size())
      return set_error(B200_ERR_INVALID, "TMA pipeline: CTA %d finished with %zu producer ops never issued", cta, rec.ops.size() - pi);
    return 0;
  }

  // whole launch; kTmaSkip if the shape is not served (the caller then walks the gather kernel's grid)
  template <class P, int recon, bool dagger, bool xpay, OpType op>
  int tma_emu_launch(const LaunchRequest &rq, const DslashArgs<P, recon> &arg)
  {
    if constexpr (P::fixed || op == OP_TM || op == OP_TM_PC || op == OP_TM_PC_PRE) {
      return kTmaSkip;
    } else {
      if (op == OP_CLOVER_PC && dagger && xpay) return kTmaSkip;
      TmaPlan plan;
      if (!tma_make_plan<P, recon>(plan, arg.geom, arg.n_parity, arg.parity, rq.tma_ty, rq.tma_tz)) return kTmaSkip;
      if (rq.tma_link_slots >= 2 && rq.tma_link_slots < plan.n_link_slots) {
        plan.n_link_slots = rq.tma_link_slots;
        plan.off_bar = plan.off_link + plan.n_link_slots * plan.link_bytes;
        plan.smem_bytes = plan.off_bar + 1024;
      }
      if (plan.smem_bytes > kTmaSmemBudget) return set_error(B200_ERR_INVALID, "TMA plan needs %d bytes of shared memory", plan.smem_bytes);
      TmaDesc descs[2][TM_COUNT];
      for (int pi = 0; pi < arg.n_parity; pi++) {
        tma_make_descs(descs[pi], arg, plan, arg.n_parity == 2 ? pi : arg.parity);
        for (int k = 0; k < TM_COUNT; k++)
          if (!tma_desc_ok(descs[pi][k])) return kTmaSkip;
      }
      int grid = rq.tma_grid > 0 ? rq.tma_grid : 148;
      if (grid > plan.n_items) grid = plan.n_items;
      long visited = 0;
      int err = 0;
#pragma omp parallel for reduction(+ : visited) schedule(dynamic)
      for (int cta = 0; cta < grid; cta++) {
        long v = 0;
        const int e = tma_emu_cta<P, recon, dagger, xpay, op>(arg, plan, descs, cta, grid, v);
        visited += v;
        if (e) {
#pragma omp critical
          err = e;
        }
      }
      if (err) return err;
      if (visited != (long)arg.geom.volume_cb * arg.n_parity)
        return set_error(B200_ERR_INVALID, "TMA work ranges visited %ld of %ld sites", visited, (long)arg.geom.volume_cb * arg.n_parity);
      return 0;
    }
  }

} // namespace b200
