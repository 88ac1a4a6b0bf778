// xsph.cu — OPT-IN extension: XSPH viscosity and vorticity confinement (PBF_FLAG_XSPH / PBF_FLAG_VORTICITY).
//
// No reference backend has either term: src/sph_constants.h:13-14 declares `C` and `VORTICITY_EPSILON`, nothing reads
// them, and finalise (ompsph.hpp:256-264) is just v = (dx/dt + v) * VD.  BASELINE's north star names
// both, so they exist here, default OFF, defined as Macklin & Mueller 2013 §5 on the step's final state and pinned by
// oracle/pbf_oracle.c (PBF_ORACLE_XSPH / _VORTICITY), which these kernels follow operation for operation:
//   positions pStar (scaled units) after the last solver iteration, the velocities finalise has just produced, the
//   step's cell table, W = poly6Kernel, grad W = spikyKernelGradient (ompsph.hpp:67-75), 27-cell visiting order, Jacobi:
//     omega_i = (1/RHO) sum_j (v_j - v_i) x grad_{p_j} W(p_i - p_j)          grad_{p_j} W = -spiky(p_i, p_j)
//     eta_i   = sum_j |omega_j| spiky(p_i, p_j),   N = eta / |eta|            (no force when |eta| < EPSILON)
//     v_i    += dt * VORTICITY_EPSILON * (N x omega_i)  +  C * sum_j (v_j - v_i) poly6(|p_i - p_j|)
// Two one-pass neighbour kernels (the passes run once per step, not once per iteration).
#include "cells.cuh"
#include "common.cuh"
#include "pair_math.cuh"

namespace pbf {

namespace {

constexpr int kBlock = 128;
constexpr float kC_XSPH = 0.00001f;            // sph_constants.h:13
constexpr float kVORTICITY_EPSILON = 0.0005f;  // sph_constants.h:14

__global__ void __launch_bounds__(kBlock) vorticity_kernel(StepConst c, const uint32_t *__restrict__ keys,
                                                           const uint32_t *__restrict__ table,
                                                           const float4 *__restrict__ pstar,
                                                           const float4 *__restrict__ vel, float4 *__restrict__ omega) {
  const uint32_t a = blockIdx.x * kBlock + threadIdx.x;
  if (a >= c.n) return;
  const float4 pa = ldg4(pstar + a), va = ldg4(vel + a);
  float wx = 0.f, wy = 0.f, wz = 0.f;
  for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) {
    const float4 pb = ldg4(pstar + b);
    const float r = strict_distance(pa, pb);
    if (!(r >= kEPSILON && r <= c.h)) return;
    const float sc = strict_spiky(r, c);
    const float gx = -fmul(fsub(pa.x, pb.x), sc), gy = -fmul(fsub(pa.y, pb.y), sc), gz = -fmul(fsub(pa.z, pb.z), sc);
    const float4 vb = ldg4(vel + b);
    const float dx = fsub(vb.x, va.x), dy = fsub(vb.y, va.y), dz = fsub(vb.z, va.z);
    wx = fadd(wx, fsub(fmul(dy, gz), fmul(dz, gy)));
    wy = fadd(wy, fsub(fmul(dz, gx), fmul(dx, gz)));
    wz = fadd(wz, fsub(fmul(dx, gy), fmul(dy, gx)));
  });
  wx = fmul(wx, kRHO_RECIP); wy = fmul(wy, kRHO_RECIP); wz = fmul(wz, kRHO_RECIP);
  omega[a] = make_float4(wx, wy, wz, fsqrt(fadd(fadd(fmul(wx, wx), fmul(wy, wy)), fmul(wz, wz))));
}

template <bool kXsph, bool kVort>
__global__ void __launch_bounds__(kBlock) xsph_vorticity_apply_kernel(StepConst c, const uint32_t *__restrict__ keys,
                                                                      const uint32_t *__restrict__ table,
                                                                      const float4 *__restrict__ pstar,
                                                                      const float4 *__restrict__ vel,
                                                                      const float4 *__restrict__ omega,
                                                                      float4 *__restrict__ vel_out) {
  const uint32_t a = blockIdx.x * kBlock + threadIdx.x;
  if (a >= c.n) return;
  const float4 pa = ldg4(pstar + a), va = ldg4(vel + a);
  float ex = 0.f, ey = 0.f, ez = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
  for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) {
    const float4 pb = ldg4(pstar + b);
    const float r = strict_distance(pa, pb);
    if (r > c.h) return;
    if (kXsph) {
      const float w = strict_poly6(r, c);
      const float4 vb = ldg4(vel + b);
      sx = fadd(sx, fmul(fsub(vb.x, va.x), w));
      sy = fadd(sy, fmul(fsub(vb.y, va.y), w));
      sz = fadd(sz, fmul(fsub(vb.z, va.z), w));
    }
    if (kVort && r >= kEPSILON) {
      const float sc = fmul(strict_spiky(r, c), __ldg(&omega[b].w));
      ex = fadd(ex, fmul(fsub(pa.x, pb.x), sc));
      ey = fadd(ey, fmul(fsub(pa.y, pb.y), sc));
      ez = fadd(ez, fmul(fsub(pa.z, pb.z), sc));
    }
  });
  float ox = va.x, oy = va.y, oz = va.z;
  if (kVort) {
    const float len = fsqrt(fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez)));
    if (len >= kEPSILON) {
      const float nx = fdiv(ex, len), ny = fdiv(ey, len), nz = fdiv(ez, len);
      const float4 w = ldg4(omega + a);
      const float k = fmul(c.dt, kVORTICITY_EPSILON);
      ox = fadd(ox, fmul(k, fsub(fmul(ny, w.z), fmul(nz, w.y))));
      oy = fadd(oy, fmul(k, fsub(fmul(nz, w.x), fmul(nx, w.z))));
      oz = fadd(oz, fmul(k, fsub(fmul(nx, w.y), fmul(ny, w.x))));
    }
  }
  if (kXsph) {
    ox = fadd(ox, fmul(kC_XSPH, sx));
    oy = fadd(oy, fmul(kC_XSPH, sy));
    oz = fadd(oz, fmul(kC_XSPH, sz));
  }
  vel_out[a] = make_float4(ox, oy, oz, va.w);
}

}  // namespace

// After finalise: vel <- vel + vorticity confinement + XSPH viscosity, per ctx->flags.  scratch_omega / scratch_vel hold n
// float4 each (the solver's pStar ping-pong buffer and the spare position buffer are free at this point).
int launch_xsph_vorticity(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *pstar, float4 *vel,
                          float4 *scratch_omega, float4 *scratch_vel) {
  const bool xs = (ctx->flags & PBF_FLAG_XSPH) != 0, vo = (ctx->flags & PBF_FLAG_VORTICITY) != 0;
  if (!xs && !vo) return PBF_OK;
  PhaseScope ps(ctx, PBF_PH_FINALISE);
  const uint32_t n = ctx->sc.n;
  const unsigned blocks = div_up(n, kBlock);
  if (vo) {
    vorticity_kernel<<<blocks, kBlock, 0, ctx->stream>>>(ctx->sc, keys_sorted, table, pstar, vel, scratch_omega);
    PBF_LAUNCH_CHECK(ctx);
  }
  if (xs && vo)
    xsph_vorticity_apply_kernel<true, true><<<blocks, kBlock, 0, ctx->stream>>>(ctx->sc, keys_sorted, table, pstar, vel,
                                                                                scratch_omega, scratch_vel);
  else if (xs)
    xsph_vorticity_apply_kernel<true, false><<<blocks, kBlock, 0, ctx->stream>>>(ctx->sc, keys_sorted, table, pstar, vel,
                                                                                 scratch_omega, scratch_vel);
  else
    xsph_vorticity_apply_kernel<false, true><<<blocks, kBlock, 0, ctx->stream>>>(ctx->sc, keys_sorted, table, pstar, vel,
                                                                                 scratch_omega, scratch_vel);
  PBF_LAUNCH_CHECK(ctx);
  PBF_CUDA(ctx, cudaMemcpyAsync(vel, scratch_vel, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  return PBF_OK;
}

}  // namespace pbf
