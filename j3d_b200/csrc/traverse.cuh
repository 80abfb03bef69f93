// traverse.cuh — device helpers shared by the traversal kernels (cast.cu, allhits.cu): the box test of the
// 8-wide quantised node (common.cuh, WideNode) and small ray utilities.
#pragma once
#include "common.cuh"

static __device__ __forceinline__ float pick(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

static __device__ __forceinline__ float safe_rcp(float d) {
  // box tests only: keep the reciprocal finite so 0 * inf never produces NaN
  const float big = 1e18f;
  if (fabsf(d) < 1e-18f) return d < 0.f ? -big : big;
  return 1.f / d;
}

// Quantised plane byte -> float without an I2F (quarter-rate XU pipe) and without any arithmetic: one PRMT
// assembles the bits 0x3F80_qq_00 = 1 + q * 2^-15 from the child's 8-byte box record, whose bytes 6 and 7 hold
// 0x80 and 0x3F (selector nibble 0xF = byte 7 with sign replication = 0x00).  The node stores its quantisation
// step pre-multiplied by 2^15, so  t = fma(m, S, B)  with  S = step * 2^15 / d  and  B = (origin - o) / d - S
// equals q * step / d + (origin - o) / d.  The cancellation of S costs at most |S| * 2^-24 (1/512 of a quantum);
// B is widened by |S| * 2^-22 on the near and far side to stay conservative.
static __device__ __forceinline__ float plane(uint32_t lo, uint32_t hi, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return __uint_as_float(r);
}
static __device__ __forceinline__ uint32_t plane_sel(uint32_t byte_index) { return 0x760Fu | (byte_index << 4); }

// Per node and axis: S and the widened near / far offsets.
struct Slab { float S, Bn, Bf; };
static __device__ __forceinline__ Slab slab(float step32k, float origin, float o, float inv_d) {
  Slab r;
  r.S = step32k * inv_d;
  const float B = fmaf(origin - o, inv_d, -r.S);
  const float pad = fabsf(r.S) * 2.384185791015625e-07f;  // 2^-22
  r.Bn = B - pad;
  r.Bf = B + pad;
  return r;
}


// 256-bit read-only load (sm_100: LDG.E.256): a 128-byte node is four of these instead of eight 128-bit loads, which
// halves the L1 tag look-ups of a warp whose lanes sit at 32 different nodes.  p must be 32-byte aligned.
struct __align__(32) U32x8 { uint32_t v[8]; };
static __device__ __forceinline__ U32x8 ldg256(const void* p) {
  U32x8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}
