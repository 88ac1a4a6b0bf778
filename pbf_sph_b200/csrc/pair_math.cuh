// pair_math.cuh — per-pair arithmetic of the lambda and delta passes, shared by the global and tiled kernels.
//
// kStrict = true : every operation is the oracle's operation in the oracle's order (oracle/pbf_oracle.c, itself
//                  bit-identical to ompsph.hpp:67-75,217-248 on the reference): no FMA contraction, IEEE divide
//                  and sqrt, the r <= h tests on the rounded sqrt.  The only non-identical piece is pow(x,4):
//                  glibc powf is not correctly rounded (<= 0.82 ulp), here x^4 is formed in double and rounded.
// kStrict = false: the production arithmetic.  Same formulas and the same summation order, but FMA contraction is
//                  allowed, r comes from rsqrt, divisions by constants are multiplications by their reciprocals and
//                  constant factors are hoisted out of the sums.  Differences stay at the 1e-7 relative level per
//                  pair; tests/test_parity_gpu.py holds the resulting step within 1e-5 of the domain size.
#pragma once

#include "common.cuh"

namespace pbf {

// 1/sqrt(x) for normal x > 0 in one MUFU (no denormal pre-scaling: every caller guarantees x >= r2_min ~ 1e-16)
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed pairs of floats (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE single operations per instruction) -----------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ float strict_distance(const float4 a, const float4 b) {
  const float dx = fsub(b.x, a.x), dy = fsub(b.y, a.y), dz = fsub(b.z, a.z);  // glm::distance = length(b - a)
  return fsqrt(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));          // dot = (x*x + y*y) + z*z
}

// poly6Kernel — ompsph.hpp:67-69
__device__ __forceinline__ float strict_poly6(float r, const StepConst &c) {
  if (r <= c.h) {
    const float d = fsub(fmul(c.h, c.h), fmul(r, r));
    return fmul(c.P6, fmul(fmul(d, d), d));
  }
  return 0.f;
}
// scalar part of spikyKernelGradient — ompsph.hpp:71-75
__device__ __forceinline__ float strict_spiky(float r, const StepConst &c) {
  const float hr = fsub(c.h, r);
  return fmul(c.SP, fdiv(fmul(hr, hr), r));
}

template <bool kStrict> struct LambdaAcc;
template <bool kStrict> struct DeltaAcc;

// ------------------------------------------------------------------------------------------ lambda, strict
template <> struct LambdaAcc<true> {
  float gx, gy, gz, rho_unit;  // rho is accumulated as mass * poly6 per pair; mass is needed per pair
  float mass;
  __device__ __forceinline__ void init() { gx = gy = gz = 0.f; rho_unit = 0.f; mass = 1.f; }
  __device__ __forceinline__ void add(const StepConst &c, const float4 pa, const float4 pb) {
    const float r = strict_distance(pa, pb);
    if (r <= c.h) {
      if (r >= kEPSILON) {
        const float s = strict_spiky(r, c);
        gx = fadd(gx, fmul(fmul(fsub(pa.x, pb.x), s), kRHO_RECIP));
        gy = fadd(gy, fmul(fmul(fsub(pa.y, pb.y), s), kRHO_RECIP));
        gz = fadd(gz, fmul(fmul(fsub(pa.z, pb.z), s), kRHO_RECIP));
      }
      rho_unit = fadd(rho_unit, fmul(mass, strict_poly6(r, c)));
    }
  }
  // strict accumulation needs the particle's mass before the loop
  __device__ __forceinline__ void set_mass(float m) { mass = m; }
  // in-radius test on r^2: sqrt is monotone, so fsqrt(r2) <= h  <=>  r2 <= r2_max (the same neighbour set)
  static __device__ __forceinline__ bool test(const StepConst &c, const float4 pa, const float4 pb) {
    const float dx = fsub(pb.x, pa.x), dy = fsub(pb.y, pa.y), dz = fsub(pb.z, pa.z);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)) <= c.r2_max;
  }
  __device__ __forceinline__ void add_in(const StepConst &c, const float4 pa, const float4 pb) { add(c, pa, pb); }
  __device__ __forceinline__ float finish(const StepConst &c, float /*mass*/, float &rho) {
    rho = rho_unit;
    const float norm2 = fadd(fadd(fmul(gx, gx), fmul(gy, gy)), fmul(gz, gz));
    const float Ci = fsub(fdiv(rho, kRHO), 1.0f);
    return fdiv(-Ci, fadd(norm2, kCFM_EPSILON));
  }
};

// ------------------------------------------------------------------------------------------ lambda, fast
template <> struct LambdaAcc<false> {
  float gx, gy, gz, t3;  // sum of d*(h-r)^2/r  and  sum of (h^2-r^2)^3
  __device__ __forceinline__ void init() { gx = gy = gz = 0.f; t3 = 0.f; }
  __device__ __forceinline__ void set_mass(float) {}
  __device__ __forceinline__ void add(const StepConst &c, const float4 pa, const float4 pb) {
    const float dx = pa.x - pb.x, dy = pa.y - pb.y, dz = pa.z - pb.z;
    const float r2 = dx * dx + dy * dy + dz * dz;
    if (r2 <= c.r2_max) {
      const float t = c.h2 - r2;
      t3 += t * t * t;
      if (r2 >= c.r2_min) {
        const float rinv = rsqrt_fast(r2);
        const float hr = c.h - r2 * rinv;
        const float s = hr * hr * rinv;
        gx += dx * s; gy += dy * s; gz += dz * s;
      }
    }
  }
  static __device__ __forceinline__ bool test(const StepConst &c, const float4 pa, const float4 pb) {
    const float dx = pa.x - pb.x, dy = pa.y - pb.y, dz = pa.z - pb.z;
    return dx * dx + dy * dy + dz * dz <= c.r2_max;
  }
  // pair already known to satisfy test()
  __device__ __forceinline__ void add_in(const StepConst &c, const float4 pa, const float4 pb) {
    const float dx = pa.x - pb.x, dy = pa.y - pb.y, dz = pa.z - pb.z;
    const float r2 = dx * dx + dy * dy + dz * dz;
    const float t = c.h2 - r2;
    t3 += t * t * t;
    if (r2 >= c.r2_min) {
      const float rinv = rsqrt_fast(r2);
      const float hr = c.h - r2 * rinv;
      const float s = hr * hr * rinv;
      gx += dx * s; gy += dy * s; gz += dz * s;
    }
  }
  __device__ __forceinline__ float finish(const StepConst &c, float mass, float &rho) {
    rho = mass * c.P6 * t3;
    const float k = c.sp_rho;
    const float ax = gx * k, ay = gy * k, az = gz * k;
    const float norm2 = ax * ax + ay * ay + az * az;
    const float Ci = rho * c.inv_rho - 1.0f;
    return -Ci / (norm2 + kCFM_EPSILON);
  }
};

// ------------------------------------------------------------------------------------------ delta, strict
template <> struct DeltaAcc<true> {
  float dx, dy, dz;
  __device__ __forceinline__ void init() { dx = dy = dz = 0.f; }
  __device__ __forceinline__ void add(const StepConst &c, const float4 pa, const float4 pb) {
    const float r = strict_distance(pa, pb);
    if (r >= kEPSILON && r <= c.h) {
      const double q = (double)fdiv(strict_poly6(r, c), c.P6dq);
      const double q2 = q * q;
      const float corr = fmul(-kCorrK, (float)(q2 * q2));  // glm::pow(x, CorrN), CorrN = 4
      const float factor = fdiv(fadd(fadd(pa.w, pb.w), corr), kRHO);
      const float s = strict_spiky(r, c);
      dx = fadd(dx, fmul(fmul(fsub(pa.x, pb.x), s), factor));
      dy = fadd(dy, fmul(fmul(fsub(pa.y, pb.y), s), factor));
      dz = fadd(dz, fmul(fmul(fsub(pa.z, pb.z), s), factor));
    }
  }
  static __device__ __forceinline__ bool test(const StepConst &c, const float4 pa, const float4 pb) {
    const float ex = fsub(pb.x, pa.x), ey = fsub(pb.y, pa.y), ez = fsub(pb.z, pa.z);
    const float r2 = fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez));
    return r2 <= c.r2_max && r2 >= c.r2_min;
  }
  __device__ __forceinline__ void add_in(const StepConst &c, const float4 pa, const float4 pb) { add(c, pa, pb); }
  __device__ __forceinline__ float4 finish(const StepConst &c, const float4 pa) {
    return clamp_to_box(c, pa.x, pa.y, pa.z, dx, dy, dz);
  }
  // pos = (pStar + deltaP) * scale; clamp to [minBound, maxBound]; pStar = pos / scale — ompsph.hpp:245-247
  static __device__ __forceinline__ float4 clamp_to_box(const StepConst &c, float px, float py, float pz, float ax,
                                                        float ay, float az) {
    float o[3];
    const float p[3] = {px, py, pz}, d[3] = {ax, ay, az};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = fmul(fadd(p[k], d[k]), c.scale);
      v = glm_min(c.max_bound[k], glm_max(c.min_bound[k], v));
      o[k] = fdiv(v, c.scale);
    }
    return make_float4(o[0], o[1], o[2], 0.f);
  }
};

// ------------------------------------------------------------------------------------------ delta, fast
template <> struct DeltaAcc<false> {
  float dx, dy, dz;
  __device__ __forceinline__ void init() { dx = dy = dz = 0.f; }
  __device__ __forceinline__ void add(const StepConst &c, const float4 pa, const float4 pb) {
    const float ex = pa.x - pb.x, ey = pa.y - pb.y, ez = pa.z - pb.z;
    const float r2 = ex * ex + ey * ey + ez * ez;
    if (r2 <= c.r2_max && r2 >= c.r2_min) {
      const float rinv = rsqrt_fast(r2);
      const float hr = c.h - r2 * rinv;
      const float t = c.h2 - r2;
      const float q = c.p6_over_dq * (t * t * t);
      const float q2 = q * q;
      const float lam = (pa.w + pb.w) - kCorrK * (q2 * q2);
      const float s = (hr * hr * rinv) * lam;
      dx += ex * s; dy += ey * s; dz += ez * s;
    }
  }
  static __device__ __forceinline__ bool test(const StepConst &c, const float4 pa, const float4 pb) {
    const float ex = pa.x - pb.x, ey = pa.y - pb.y, ez = pa.z - pb.z;
    const float r2 = ex * ex + ey * ey + ez * ez;
    return r2 <= c.r2_max && r2 >= c.r2_min;
  }
  __device__ __forceinline__ void add_in(const StepConst &c, const float4 pa, const float4 pb) {
    const float ex = pa.x - pb.x, ey = pa.y - pb.y, ez = pa.z - pb.z;
    const float r2 = ex * ex + ey * ey + ez * ez;
    if (r2 >= c.r2_min) {  // the lambda pass's hit list contains the particle itself (r = 0)
      const float rinv = rsqrt_fast(r2);
      const float hr = c.h - r2 * rinv;
      const float t = c.h2 - r2;
      const float q = c.p6_over_dq * (t * t * t);
      const float q2 = q * q;
      const float lam = (pa.w + pb.w) - kCorrK * (q2 * q2);
      const float s = (hr * hr * rinv) * lam;
      dx += ex * s; dy += ey * s; dz += ez * s;
    }
  }
  __device__ __forceinline__ float4 finish(const StepConst &c, const float4 pa) {
    const float k = c.SP * c.inv_rho;
    return DeltaAcc<true>::clamp_to_box(c, pa.x, pa.y, pa.z, dx * k, dy * k, dz * k);
  }
};

}  // namespace pbf
