// PTX helpers of the mma.sync / TMA attention kernels (attn_tiles.cu, attn_long.cu): mbarrier + TMA tensor copies,
// ldmatrix / mma.sync / movmatrix wrappers and the swizzled-tile addressing both files share.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace amc {
// bf16 tensor [B][T][cols] (row pitch = cols) as a 3-D TMA map, box = {dh, box_rows, 1}, swizzle span = dh * 2 bytes
int attn_make_map3(CUtensorMap* map, const void* base, int B, int T, int cols, int dh, int box_rows);
int attn_sm_count();

namespace attn_ptx {

// ---- PTX helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t a) {   // 8x8 b16 transpose across the warp
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64f(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ---- swizzled tile addressing --------------------------------------------------------------------------------
// A tile is Tpad rows of RB = 32*KD bytes written by TMA with SWIZZLE_{32,64,128}B: the 16-byte chunk index is
// XORed with address bits [7, 7+log2(chunks per row)); tile bases are aligned to 1024 B so those bits are a
// function of the row alone.
template <int KD> __device__ __forceinline__ int swz(int r) {
  return KD == 1 ? ((r >> 2) & 1) : (KD == 2 ? ((r >> 1) & 3) : (r & 7));
}
template <int KD> __device__ __forceinline__ uint32_t chunk_addr(uint32_t tile, int r, int ch) {
  return tile + (uint32_t)(r * (32 * KD)) + (uint32_t)((ch ^ swz<KD>(r)) << 4);
}
// ldmatrix.x4 lane addresses over a 16-row x 16-column block (rows row0.., columns 16*ks..):
//  pattern A: matrices = (rows 0-7, cols 0-7), (rows 8-15, cols 0-7), (rows 0-7, cols 8-15), (rows 8-15, cols 8-15)
//             -> A operand (non-trans) / [k][n] B operand of two n-blocks (with .trans)
//  pattern B: matrices = (rows 0-7, cols 0-7), (rows 0-7, cols 8-15), (rows 8-15, cols 0-7), (rows 8-15, cols 8-15)
//             -> B operand pairs {b0,b1} for n-block rows 0-7 and {b2,b3} for n-block rows 8-15 (rows = n, cols = k)
template <int KD> __device__ __forceinline__ uint32_t addrA(uint32_t tile, int row0, int ks, int lane) {
  return chunk_addr<KD>(tile, row0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4));
}
template <int KD> __device__ __forceinline__ uint32_t addrB(uint32_t tile, int row0, int ks, int lane) {
  return chunk_addr<KD>(tile, row0 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1));
}

// ---- forward building block --------------------------------------------------------------------------------
// One chunk of NBC key blocks (16 keys each) of a 16-query-row item: S = Q K^T, online-softmax update, O += P V.
// LAST = false: all NBC blocks exist and none needs masking -> straight-line code.
// LAST = true : the final chunk of the row: `nblk` (1..NBC) blocks exist and the zero-filled key rows >= T of the
//               very last block are masked to -inf.
template <int KD, int NBC, bool LAST>
__device__ __forceinline__ void fwd_chunk(uint32_t kb, uint32_t vb, int k0, int nblk, int T, const uint32_t (&aq)[KD][4],
                                          float (&o)[2 * KD][4], float& m0, float& m1, float& l0, float& l1, float sl2,
                                          int lane) {
  const int cb = (lane & 3) * 2;
  float c[2 * NBC][4];
#pragma unroll
  for (int n = 0; n < 2 * NBC; ++n) { c[n][0] = 0.f; c[n][1] = 0.f; c[n][2] = 0.f; c[n][3] = 0.f; }
#pragma unroll
  for (int j = 0; j < NBC; ++j) {
    if (!LAST || j < nblk) {
#pragma unroll
      for (int ks = 0; ks < KD; ++ks) {
        uint32_t bfr[4];
        ldsm_x4(bfr, addrB<KD>(kb, (k0 + j) * 16, ks, lane));
        mma_bf16(c[2 * j], aq[ks], bfr[0], bfr[1]);
        mma_bf16(c[2 * j + 1], aq[ks], bfr[2], bfr[3]);
      }
    }
  }
  if (LAST) {      // keys >= T (zero-filled rows of the last block, and blocks past the end) take no probability mass
#pragma unroll
    for (int n = 0; n < 2 * NBC; ++n) {
      const int key = k0 * 16 + n * 8 + cb;
      if (key >= T) { c[n][0] = -INFINITY; c[n][2] = -INFINITY; }
      if (key + 1 >= T) { c[n][1] = -INFINITY; c[n][3] = -INFINITY; }
    }
  }
  float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < 2 * NBC; ++n) {
    x0 = fmaxf(x0, fmaxf(c[n][0], c[n][1]));
    x1 = fmaxf(x1, fmaxf(c[n][2], c[n][3]));
  }
  x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
  x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
  const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);
  const float al0 = ex2((m0 - n0) * sl2), al1 = ex2((m1 - n1) * sl2);   // first chunk: ex2(-inf) = 0
  m0 = n0; m1 = n1;
  const float ms0 = m0 * sl2, ms1 = m1 * sl2;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < 2 * NBC; ++n) {
    c[n][0] = ex2(fmaf(c[n][0], sl2, -ms0)); c[n][1] = ex2(fmaf(c[n][1], sl2, -ms0));
    c[n][2] = ex2(fmaf(c[n][2], sl2, -ms1)); c[n][3] = ex2(fmaf(c[n][3], sl2, -ms1));
    s0 += c[n][0] + c[n][1];
    s1 += c[n][2] + c[n][3];
  }
  l0 = fmaf(l0, al0, s0);
  l1 = fmaf(l1, al1, s1);
#pragma unroll
  for (int n = 0; n < 2 * KD; ++n) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }
#pragma unroll
  for (int j = 0; j < NBC; ++j) {
    if (!LAST || j < nblk) {
      const uint32_t pa[4] = {pack2(c[2 * j][0], c[2 * j][1]), pack2(c[2 * j][2], c[2 * j][3]),
                              pack2(c[2 * j + 1][0], c[2 * j + 1][1]), pack2(c[2 * j + 1][2], c[2 * j + 1][3])};
#pragma unroll
      for (int np = 0; np < KD; ++np) {
        uint32_t bfr[4];
        ldsm_x4_t(bfr, addrA<KD>(vb, (k0 + j) * 16, np, lane));
        mma_bf16(o[2 * np], pa, bfr[0], bfr[1]);
        mma_bf16(o[2 * np + 1], pa, bfr[2], bfr[3]);
      }
    }
  }
}

}  // namespace attn_ptx
}  // namespace amc
