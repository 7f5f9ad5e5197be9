// Device-side building blocks shared by the fused RHS kernel (hg_fused.cu) and the fused VJP kernel
// (hg_vjp.cu): TMA bulk-copy + mbarrier wrappers, branch-free fp64 helpers, the per-side state record and
// the face-once Roe flux.
#pragma once
#include "hg_ctx.h"

namespace hg {
namespace dev {

// ---------------------------------------------------------------- TMA bulk copy + mbarrier (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// raise the barrier's pending byte count WITHOUT arriving: lets a thread start copies before it knows the total
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of a contiguous block (16-byte aligned, size a multiple of 16): nothing is written on the SM
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// L2 prefetch of [p, p + bytes) by one warp through the load/store path (prefetch.global.L2, one 64-byte chunk per lane and
// trip).  The copy engine's cp.async.bulk.prefetch.L2 would take one of its operation slots per row, and that engine -- one
// per SM, ~100 cycles per operation under load -- is what the tile kernels run out of first.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void warp_prefetch_l2(const void* p, uint32_t bytes, int lane) {
  const char* c = static_cast<const char*>(p);
  for (uint32_t off = (uint32_t)lane * 64u; off < bytes + 64u; off += 2048u) prefetch_l2(c + min(off, bytes - 1u));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The tile kernels are launched as programmatic dependents of the small single-CTA kernel that precedes them
// (k_inlet_coef: the boundary-wide conveyance sum): it releases its dependents at once, the tile kernel runs beside it, and
// only the threads that evaluate an inlet-q face wait for its result.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- library-owned halo exchange: consumer side
// Thread k < w.n of a band tile's CTA waits until neighbour k has published this exchange's epoch (hg_comm.cu); the
// block barrier that follows phase 1 hands the acquired view to the rest of the CTA.  Gives up after ~4 s (a peer that
// never pushed: mismatched call sequences) by raising the device error flag instead of hanging the GPU.
__device__ __forceinline__ void comm_wait(const CommWait& w, int tid) {
  if (tid >= w.n) return;
  const unsigned long long* f = w.flags + (size_t)tid * 16;
  unsigned long long v, t0 = 0, t;
  for (uint32_t it = 0;; ++it) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v >= w.epoch) return;
    if ((it & 255u) == 255u) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) { atomicExch(w.err, HG_ERR_COMM); return; }
    }
    __nanosleep(64);
  }
}

// ---------------------------------------------------------------- branch-free fp64 helpers
// Every argument on this path is a positive normal number (h >= h_small, x^2 + eps, g h + eps, areas,
// Manning's n), so the IEEE special-case slow paths of '/', sqrt() and cbrt() are dead weight: each costs
// a branch + call sequence per use.  These are the same MUFU seed + Newton refinements, straight-line;
// results are within 1-2 ulp of the correctly rounded value (parity budget is 1e-12).
// The MUFU seeds read only the high word of the argument (MUFU.RCP64H / RSQ64H: ~2^-20 relative error), so ONE third-order
// step (error -> ~e^3 = 2^-60) reaches full precision where two quadratic steps (2^-40, 2^-80) were needed: fewer fp64-pipe
// instructions and a shorter dependency chain.  Accuracy is measured on the device by tests/test_gpu_parity.py::test_fast_math.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));   // MUFU.RCP64H
  const double e = fma(-x, r, 1.0);                        // 1/x = r / (1 - e) = r (1 + e + e^2 + ...)
  return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); // MUFU.RSQ64H
  const double e = fma(-(x * r), r, 1.0);                  // x^(-1/2) = r (1 - e)^(-1/2) = r (1 + e/2 + 3 e^2/8 + ...)
  return fma(r * e, fma(0.375, e, 0.5), r);
}
__device__ __forceinline__ double fast_sqrt(double x) {   // x > 0
  const double r = fast_rsqrt(x);
  const double s = x * r;
  return fma(fma(-s, s, x), 0.5 * r, s);                   // one correction: correctly rounded but for rare ties
}
// w = x^(-1/3) for x > 0: float seed (~2^-21) + ONE third-order step, w <- w (1 + e/3 + 2 e^2/9), e = 1 - x w^3.
// x^(-7/3) = w^7 (friction, semi_discretize_swe_2D.jl:544-547) and 1/x = w^3 both come from it.
__device__ __forceinline__ double rcbrt_pos(double x) {
  const double w = (double)exp2f(-0.33333334f * log2f((float)x));
  const double e = fma(-x, w * w * w, 1.0);
  return fma(w * e, fma(0.2222222222222222, e, 0.3333333333333333), w);
}
__device__ __forceinline__ double pow_m73(double x) {
  const double w = rcbrt_pos(x);
  const double w2 = w * w, w4 = w2 * w2;
  return w4 * w2 * w;
}
// smooth |x| = sqrt(x^2 + eps) (utilities/smooth_functions.jl), three per face: y * rsqrt(y) without the correction step,
// <= 1 ulp (rounding of r and of the product)
__device__ __forceinline__ double smooth_abs(double x) {
  const double y = fma(x, x, EPS);
  return y * fast_rsqrt(y);
}

struct Side {
  double xi, h, hu, hv, zb, u, v, s, P;
};

// Riemann_2D_Roe, face-once form.  Returns (flux along the face normal, outward for L) * len.  The 1/2 of the Roe average
// of the two physical fluxes rides in the length factor (scaling by 2 is exact, so the bits are those of 0.5*(..)*len).
// zbL() / zbR() fetch the bed elevations; they are only called on the (rare) faces with a dry side, so the bed is never
// staged: it stays in global memory.
template <class ZL, class ZR>
__device__ __forceinline__ void roe_flux(Side L, Side R, ZL zbL, ZR zbR, double nx, double ny, double len,
                                         double g, double hmin, double& o0, double& o1, double& o2) {
  const bool dryL = L.h <= hmin, dryR = R.h <= hmin;
  if (dryL || dryR) {
    if (dryL && dryR) { o0 = o1 = o2 = 0.0; return; }         // swe_2D_solvers.jl:16
    L.zb = zbL(); R.zb = zbR();
    if ((L.h + L.zb) < (R.zb + hmin) && dryR) {                // :23 wall-like: mirror L into R
      R.h = L.h; R.hu = -L.hu; R.hv = -L.hv; R.u = -L.u; R.v = -L.v; R.s = L.s;
    } else if ((R.h + R.zb) < (L.zb + hmin) && dryL) {         // :39
      L.h = R.h; L.hu = -R.hu; L.hv = -R.hv; L.u = -R.u; L.v = -R.v; L.s = R.s;
    } else {                                                   // :54 / :65 one-sided physical flux
      const Side& W = dryL ? R : L;
      const double hp = W.h + EPS;
      const double p = 0.5 * g * hp * hp;
      const double un = W.u * nx + W.v * ny;
      o0 = (W.hu * nx + W.hv * ny) * len;
      o1 = (W.hu * un + p * nx) * len;
      o2 = (W.hv * un + p * ny) * len;
      return;
    }
  }
  const double rs = fast_rcp(L.s + R.s);
  const double uRoe = (L.s * L.u + R.s * R.u) * rs;
  const double vRoe = (L.s * L.v + R.s * R.v) * rs;
  const double un = uRoe * nx + vRoe * ny;
  const double c2 = fma(0.5 * g, L.h + R.h, EPS);             // g hRoe + eps, hRoe = (hL + hR)/2 (:91 arithmetic mean)
  const double rc = fast_rsqrt(c2);
  const double c = c2 * rc;                                    // sqrt(g hRoe + eps)
  const double k = 0.5 * rc;                                   // 1/(2c)
  const double d1 = R.xi - L.xi, d2 = R.hu - L.hu, d3 = R.hv - L.hv;
  const double w1 = -(uRoe * ny - vRoe * nx) * d1 + ny * d2 - nx * d3;   // L_mat * dQ  (:107-118)
  const double m = k * (un * d1 - (nx * d2 + ny * d3));
  const double w2 = 0.5 * d1 + m, w3 = 0.5 * d1 - m;
  const double z1 = smooth_abs(un) * w1, z2 = smooth_abs(un - c) * w2, z3 = smooth_abs(un + c) * w3;
  const double zs = z2 + z3, zd = c * (z3 - z2);
  const double y1 = zs;                                        // R_mat * (|Lambda| w)
  const double y2 = ny * z1 + uRoe * zs + nx * zd;
  const double y3 = -nx * z1 + vRoe * zs + ny * zd;
  const double unL = L.u * nx + L.v * ny, unR = R.u * nx + R.v * ny;
  const double ps = L.P + R.P;
  const double hl = 0.5 * len;
  o0 = ((L.hu * nx + L.hv * ny) + (R.hu * nx + R.hv * ny) - y1) * hl;  // :121-133
  o1 = (L.hu * unL + R.hu * unR + ps * nx - y2) * hl;
  o2 = (L.hv * unL + R.hv * unR + ps * ny - y3) * hl;
}

__device__ __forceinline__ void derive(Side& s, double hst, double g) {
  const double rh = fast_rcp(s.h);
  s.u = s.hu * rh;
  s.v = s.hv * rh;
  s.s = fast_sqrt(s.h + EPS);
  const double xe = s.xi + EPS;
  s.P = 0.5 * g * fma(xe, xe, 2.0 * s.xi * hst);  // xi-form pressure, swe_2D_solvers.jl:122
}


// Compile-time tile configuration: every shared-memory array has a constant stride, so all smem
// accesses are [register + immediate] and the per-cell face loop is fully unrolled.
template <int T_, int ML_, int MF_, int NF_, int THREADS_, int MINB_>
struct TileCfg {
  static constexpr int T = T_, ML = ML_, MF = MF_, NF = NF_, THREADS = THREADS_, MINB = MINB_;
  static constexpr int kSmem = 16 + 8 * (6 * ML + 3 * MF + 4 * T) + 4 * MF + 2 * T * NF;
  // half as many threads as cells: every thread owns exactly two cells and ~four faces, processed pairwise
  static constexpr bool kDual = 2 * THREADS_ <= T_;
};


}  // namespace dev
}  // namespace hg

// X-macro over the compiled configurations, in priority order: (id, T, ML, MF, NF, THREADS, MINB).
// MINB CTAs/SM is what the shared-memory footprint allows; THREADS keeps MINB*THREADS*regs <= 64K.
// Measured on B200 (round 1, 4M-cell river, ms per RHS): T=256/160 thr/5 CTAs 0.174, T=192/160/6 0.176,
// T=256/192/4 0.184, T=512/384/2 0.195.  Round 2 (16M cells, after the leaner arithmetic): T=256/128 thr/5 CTAs with two
// faces / cells per trip 0.531-0.538 (sustained 0.625-0.636) vs T=256/160 0.543-0.553 (0.646-0.659); T=224 0.56, T=192 0.67.
#define HG_TILE_CONFIGS(X)            \
  X(10, 256, 336, 564, 4, 128, 5)     \
  X(7, 256, 336, 564, 4, 160, 5)      \
  X(11, 256, 336, 564, 4, 128, 4)     \
  X(1, 256, 352, 580, 4, 192, 4)      \
  X(0, 256, 352, 580, 4, 128, 4)      \
  X(2, 256, 352, 580, 4, 256, 3)      \
  X(14, 384, 480, 872, 4, 192, 3)     \
  X(4, 512, 672, 1124, 4, 384, 2)     \
  X(3, 512, 672, 1124, 4, 256, 2)     \
  X(15, 240, 320, 536, 4, 128, 5)     \
  X(12, 224, 304, 504, 4, 128, 5)     \
  X(13, 224, 304, 504, 4, 160, 5)     \
  X(9, 192, 264, 436, 4, 160, 6)      \
  X(8, 192, 264, 436, 4, 128, 6)      \
  X(6, 128, 192, 324, 4, 128, 8)      \
  X(5, 128, 256, 644, 8, 128, 4)

