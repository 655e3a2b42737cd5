// Non-GEMM pieces of the ViT-L/14 forward (row a1): im2col for the 14x14/14 patch embedding, CLS rows, LayerNorm
// (-> bf16 hi/lo planes, the A operand of the next tcgen05 GEMM), multi-head self-attention over 257 tokens, and the
// fp32 -> hi/lo plane split used to pack weights.
//
// Attention: one CTA per (crop, head); all 257 keys / values of the head are staged once in shared memory as bf16
// hi/lo planes (XOR-swizzled 16-byte chunks, conflict-free ldmatrix) and nine warps walk the 17 query row groups with
// an online softmax.  The two products use the same fp32-faithful split as the GEMMs:
//   S = Qh.Kh^T + Qh.Kl^T + Ql.Kh^T,   O = Ph.Vh + Ph.Vl + Pl.Vh      (mma.sync m16n8k16 bf16, fp32 accumulate).
// At 257 x 64 per head the problem is far below one tcgen05 tile per CTA pair and is 4 % of the ViT FLOPs; the
// warp-level MMA keeps the softmax in registers.  Softmax itself is fp32 (expf).
#include "gigapose_kernels.h"
#include <cuda_bf16.h>

namespace gp {

namespace {

constexpr int kDim = 1024, kHeads = 16, kHd = 64;
constexpr int kTok = 257;

__device__ __forceinline__ void split_store(float v, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t i) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---------------------------------------------------------------- fp32 [rows, K] -> bf16 hi/lo planes [rows, Kpad]
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int K, int Kpad,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Kpad) return;
  const long long r = i / Kpad;
  const int k = (int)(i - r * Kpad);
  split_store(k < K ? x[r * K + k] : 0.f, hi, lo, (size_t)i);
}

// ---------------------------------------------------------------- im2col: [b,3,224,224] -> planes [b*256, 608]
// column = c*196 + ky*14 + kx, the flattening of the conv weight [1024,3,14,14]; columns 588..607 are zero
__global__ void im2col_kernel(const float* __restrict__ img, int b, int Kpad, __nv_bfloat16* __restrict__ hi,
                              __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)b * 256 * Kpad;
  if (i >= total) return;
  const long long row = i / Kpad;
  const int k = (int)(i - row * Kpad);
  float v = 0.f;
  if (k < 588) {
    const int im = (int)(row >> 8), pidx = (int)(row & 255), py = pidx >> 4, px = pidx & 15;
    const int c = k / 196, rem = k - c * 196, ky = rem / 14, kx = rem - ky * 14;
    v = img[(((size_t)im * 3 + c) * 224 + (py * 14 + ky)) * 224 + px * 14 + kx];
  }
  split_store(v, hi, lo, (size_t)i);
}

// ---------------------------------------------------------------- x[b*257 + 0, :] = cls + pos[0]
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, int b, float* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b * kDim) return;
  const int im = i / kDim, c = i - im * kDim;
  x[(size_t)im * kTok * kDim + c] = cls[c] + pos[c];
}

// ---------------------------------------------------------------- LayerNorm (eps 1e-6) -> hi/lo planes; warp per row
__global__ void __launch_bounds__(256)
layernorm_planes_kernel(const float* __restrict__ x, int M, const float* __restrict__ w, const float* __restrict__ bsh,
                        float eps, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * kDim);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = src[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / kDim;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += a * a + b * b + c * c + d * d;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float rstd = rsqrtf(ss / kDim + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 wv = *reinterpret_cast<const float4*>(w + c), bv = *reinterpret_cast<const float4*>(bsh + c);
    const float o[4] = {(v[i].x - mean) * rstd * wv.x + bv.x, (v[i].y - mean) * rstd * wv.y + bv.y,
                        (v[i].z - mean) * rstd * wv.z + bv.z, (v[i].w - mean) * rstd * wv.w + bv.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(o[j]);
      l[j] = __float2bfloat16_rn(o[j] - __bfloat162float(h[j]));
    }
    const size_t off = (size_t)row * kDim + c;
    *reinterpret_cast<uint2*>(hi + off) = make_uint2((uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
                                                     (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
    *reinterpret_cast<uint2*>(lo + off) = make_uint2((uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
                                                     (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
  }
}

// ---------------------------------------------------------------- attention
constexpr int kKeyPad = 320;                       // 257 keys padded to 5 tiles of 64
constexpr int kQGroups = 17;                       // ceil(257 / 16) query row groups
constexpr int kAttnWarps = 9;                      // 17 row groups -> 9 + 8
constexpr int kPlane = kKeyPad * kHd * 2;          // 40 KB: one bf16 plane of K or V (rows of 128 B)
constexpr int kAttnSmem = 4 * kPlane + 2 * (kAttnWarps * 16 * kHd * 2);   // K hi/lo, V hi/lo + per-warp Q hi/lo

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][64 bf16] tile, XOR-swizzled
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a - __bfloat162float(ah))) |
       ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b - __bfloat162float(bh))) << 16);
}

// qkv planes [M, 3072] (q | k | v, head h at columns h*64); out planes [M, 1024]
__global__ void __launch_bounds__(kAttnWarps * 32, 1)
attention_kernel(const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo,
                 __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int passes) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int img = blockIdx.x / kHeads, head = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sKh = sm;
  uint8_t* sKl = sm + kPlane;
  uint8_t* sVh = sm + 2 * kPlane;
  uint8_t* sVl = sm + 3 * kPlane;
  uint8_t* sQh = sm + 4 * kPlane + warp * (16 * kHd * 2);
  uint8_t* sQl = sm + 4 * kPlane + kAttnWarps * (16 * kHd * 2) + warp * (16 * kHd * 2);
  const size_t row0 = (size_t)img * kTok;
  const int ld = 3 * kDim;

  // stage K and V of this head: 320 rows x 8 chunks of 16 B per plane (rows >= 257 zero)
  for (int i = threadIdx.x; i < kKeyPad * 8; i += blockDim.x) {
    const int row = i >> 3, chunk = i & 7;
    uint4 kh = make_uint4(0, 0, 0, 0), kl = kh, vh = kh, vl = kh;
    if (row < kTok) {
      const size_t g = (row0 + row) * ld + head * kHd + chunk * 8;
      kh = *reinterpret_cast<const uint4*>(qkv_hi + g + kDim);
      vh = *reinterpret_cast<const uint4*>(qkv_hi + g + 2 * kDim);
      if (passes == 3) {
        kl = *reinterpret_cast<const uint4*>(qkv_lo + g + kDim);
        vl = *reinterpret_cast<const uint4*>(qkv_lo + g + 2 * kDim);
      }
    }
    const uint32_t o = swz(row, chunk);
    *reinterpret_cast<uint4*>(sKh + o) = kh;
    *reinterpret_cast<uint4*>(sKl + o) = kl;
    *reinterpret_cast<uint4*>(sVh + o) = vh;
    *reinterpret_cast<uint4*>(sVl + o) = vl;
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  for (int qg = warp; qg < kQGroups; qg += kAttnWarps) {
    const int q0 = qg * 16;
    // stage this warp's 16 query rows (hi/lo)
    for (int i = lane; i < 16 * 8; i += 32) {
      const int row = i >> 3, chunk = i & 7;
      uint4 qh = make_uint4(0, 0, 0, 0), ql = qh;
      if (q0 + row < kTok) {
        const size_t gq = (row0 + q0 + row) * ld + head * kHd + chunk * 8;
        qh = *reinterpret_cast<const uint4*>(qkv_hi + gq);
        if (passes == 3) ql = *reinterpret_cast<const uint4*>(qkv_lo + gq);
      }
      *reinterpret_cast<uint4*>(sQh + swz(row, chunk)) = qh;
      *reinterpret_cast<uint4*>(sQl + swz(row, chunk)) = ql;
    }
    __syncwarp();
    // Q fragments for the 4 k-steps over d (A operand, 16 x 16 each)
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int row = lane & 15, chunk = ks * 2 + (lane >> 4);
      ldsm_x4((uint32_t)__cvta_generic_to_shared(sQh + swz(row, chunk)), qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
      ldsm_x4((uint32_t)__cvta_generic_to_shared(sQl + swz(row, chunk)), ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // rows g and g + 8

    for (int kt = 0; kt < kKeyPad / 64; ++kt) {
      const int key0 = kt * 64;
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
      // S = Q K^T over d: B operand = K[key][d], 8 x 8 blocks, non-transposed ldmatrix.  The three split passes are
      // issued tile-major (8 independent accumulators between two MMAs on the same tile) to hide the MMA latency.
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bh[4][4], bl[4][4];
#pragma unroll
        for (int np = 0; np < 4; ++np) {               // pairs of 8-key n-tiles
          const int row = key0 + np * 16 + (lane >> 4) * 8 + (lane & 7);
          const int chunk = ks * 2 + ((lane >> 3) & 1);
          ldsm_x4((uint32_t)__cvta_generic_to_shared(sKh + swz(row, chunk)), bh[np][0], bh[np][1], bh[np][2], bh[np][3]);
          if (passes == 3)
            ldsm_x4((uint32_t)__cvta_generic_to_shared(sKl + swz(row, chunk)), bl[np][0], bl[np][1], bl[np][2], bl[np][3]);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          mma_bf16(s[2 * np], qh[ks], bh[np][0], bh[np][1]);
          mma_bf16(s[2 * np + 1], qh[ks], bh[np][2], bh[np][3]);
        }
        if (passes == 3) {
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16(s[2 * np], qh[ks], bl[np][0], bl[np][1]);
            mma_bf16(s[2 * np + 1], qh[ks], bl[np][2], bl[np][3]);
          }
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16(s[2 * np], ql[ks], bh[np][0], bh[np][1]);
            mma_bf16(s[2 * np + 1], ql[ks], bh[np][2], bh[np][3]);
          }
        }
      }
      // scale (1/sqrt(64), exact power of two), mask padded keys, online softmax
      float mx0 = m0, mx1 = m1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int key = key0 + i * 8 + 2 * t + (j & 1);
          s[i][j] = key < kTok ? s[i][j] * 0.125f : -INFINITY;
        }
        mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
        mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float c0 = expf(m0 - mx0), c1 = expf(m1 - mx1);      // m = -inf on the first tile -> 0
      m0 = mx0; m1 = mx1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = expf(s[i][0] - m0); s[i][1] = expf(s[i][1] - m0);
        s[i][2] = expf(s[i][2] - m1); s[i][3] = expf(s[i][3] - m1);
        rs0 += s[i][0] + s[i][1];
        rs1 += s[i][2] + s[i][3];
        o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1;
      }
      l0 = l0 * c0 + rs0;
      l1 = l1 * c1 + rs1;
      // O += P V : A operand = P (from the S registers), B operand = V[key][d] via transposed ldmatrix
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                 // 16 keys per step
        uint32_t ph[4], pl[4];
        split2(s[2 * ks][0], s[2 * ks][1], ph[0], pl[0]);
        split2(s[2 * ks][2], s[2 * ks][3], ph[1], pl[1]);
        split2(s[2 * ks + 1][0], s[2 * ks + 1][1], ph[2], pl[2]);
        split2(s[2 * ks + 1][2], s[2 * ks + 1][3], ph[3], pl[3]);
        uint32_t vh[4][4], vl[4][4];
#pragma unroll
        for (int np = 0; np < 4; ++np) {               // pairs of 8-wide d n-tiles
          const int row = key0 + ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int chunk = np * 2 + (lane >> 4);
          ldsm_x4_t((uint32_t)__cvta_generic_to_shared(sVh + swz(row, chunk)), vh[np][0], vh[np][1], vh[np][2], vh[np][3]);
          if (passes == 3)
            ldsm_x4_t((uint32_t)__cvta_generic_to_shared(sVl + swz(row, chunk)), vl[np][0], vl[np][1], vl[np][2], vl[np][3]);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          mma_bf16(o[2 * np], ph, vh[np][0], vh[np][1]);
          mma_bf16(o[2 * np + 1], ph, vh[np][2], vh[np][3]);
        }
        if (passes == 3) {
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16(o[2 * np], ph, vl[np][0], vl[np][1]);
            mma_bf16(o[2 * np + 1], ph, vl[np][2], vl[np][3]);
          }
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16(o[2 * np], pl, vh[np][0], vh[np][1]);
            mma_bf16(o[2 * np + 1], pl, vh[np][2], vh[np][3]);
          }
        }
      }
    }
    // finish the row sums across the 4 lanes of a quad, normalise, store hi/lo planes
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = head * kHd + i * 8 + 2 * t;
      uint32_t h, l;
      if (r0 < kTok) {
        split2(o[i][0] * inv0, o[i][1] * inv0, h, l);
        *reinterpret_cast<uint32_t*>(out_hi + (row0 + r0) * kDim + col) = h;
        *reinterpret_cast<uint32_t*>(out_lo + (row0 + r0) * kDim + col) = l;
      }
      if (r1 < kTok) {
        split2(o[i][2] * inv1, o[i][3] * inv1, h, l);
        *reinterpret_cast<uint32_t*>(out_hi + (row0 + r1) * kDim + col) = h;
        *reinterpret_cast<uint32_t*>(out_lo + (row0 + r1) * kDim + col) = l;
      }
    }
    __syncwarp();
  }
}

}  // namespace

cudaError_t launch_split_planes(const float* x, long long rows, int K, int Kpad, uint16_t* hi, uint16_t* lo, cudaStream_t s) {
  const long long total = rows * Kpad;
  if (total <= 0) return cudaSuccess;
  split_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, rows, K, Kpad, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                     reinterpret_cast<__nv_bfloat16*>(lo));
  return cudaGetLastError();
}

cudaError_t launch_im2col(const float* img, int b, int Kpad, uint16_t* hi, uint16_t* lo, cudaStream_t s) {
  const long long total = (long long)b * 256 * Kpad;
  if (total <= 0) return cudaSuccess;
  im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(img, b, Kpad, reinterpret_cast<__nv_bfloat16*>(hi),
                                                               reinterpret_cast<__nv_bfloat16*>(lo));
  return cudaGetLastError();
}

cudaError_t launch_cls_rows(const float* cls, const float* pos, int b, float* x, cudaStream_t s) {
  if (b <= 0) return cudaSuccess;
  cls_rows_kernel<<<(b * kDim + 255) / 256, 256, 0, s>>>(cls, pos, b, x);
  return cudaGetLastError();
}

cudaError_t launch_layernorm_planes(const float* x, int M, const float* w, const float* b, float eps, uint16_t* hi,
                                    uint16_t* lo, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  layernorm_planes_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, M, w, b, eps, reinterpret_cast<__nv_bfloat16*>(hi),
                                                     reinterpret_cast<__nv_bfloat16*>(lo));
  return cudaGetLastError();
}

cudaError_t launch_attention(const uint16_t* qkv_hi, const uint16_t* qkv_lo, uint16_t* out_hi, uint16_t* out_lo, int b,
                             int passes, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (b <= 0) return cudaSuccess;
  attention_kernel<<<b * kHeads, kAttnWarps * 32, kAttnSmem, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv_hi),
                                                                 reinterpret_cast<const __nv_bfloat16*>(qkv_lo),
                                                                 reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                                 reinterpret_cast<__nv_bfloat16*>(out_lo), passes);
  return cudaGetLastError();
}

}  // namespace gp
