// Fused short-sequence attention on the 5th-generation tensor cores (tcgen05 + TMEM), bf16, head dim 16 / 32 / 64,
// 49 <= T <= 272 tokens per frame (scale_dot_product_attention.py:26-37 + the head split / concat of
// multi_head_attention.py:34-47; backward per SURVEY Appendix B).
//
// Work unit = one (frame, head).  Its [T, dh] slices of q, k, v arrive as 3-D TMA tensor copies (dims = column, token,
// frame) into swizzled shared-memory tiles (swizzle span = the row: 32 / 64 / 128 B); rows past T are zero-filled by
// the TMA unit.  The same tiles are the UMMA operands: Q and K as K-major operands of S = Q K^T, V as the MN-major B
// operand of O = P V -- nothing is transposed or re-laid-out in shared memory.
//
// Every reference sequence length is 2^k + 1 (CLS token), so a unit is split into MAIN rows -- 128-row query tiles
// (64 rows for T <= 80) that fill the TMEM lanes exactly -- and up to 16 LEFTOVER rows (the +1) that one extra warp
// runs through the mma.sync 16-row block of attn_mma.cuh against the same K / V tiles.
//
// Forward, per main tile (one CTA = RM/32 softmax warps + 1 issuer warp + 1 leftover warp, two CTAs per SM):
//   issuer thread : TMA loads (2-stage ring over units) ; S[128, Tk] = Q K^T by tcgen05.mma into TMEM ;
//                   after the softmax: O[128, dh] = P V with P read from TMEM (A operand) ; TMA store of the output
//   softmax warps : thread = query row = TMEM lane.  Pass 1 over the row (tcgen05.ld) for the max, pass 2 for
//                   p = exp2(s c - m c), the row sum and the bf16 P written back over the S columns (tcgen05.st);
//                   then O is read back, scaled by 1 / sum and staged for the TMA store (the Q tile is dead by then
//                   and serves as the staging buffer); log2-domain row statistics lse2 = m c + log2(sum) are saved.
//   No shuffles, no ldmatrix, no per-element shared-memory traffic on the main rows.
// TMEM columns: S at [0, Tk), P aliases S at [0, Tk/2) (bf16 pairs), O at [OC, OC + dh) inside the dead S columns:
// 128 / 256 columns per CTA for T <= 80 / <= 144, so 4 / 2 CTAs share an SM and one CTA's MMA and copy phases overlap
// the other's softmax.
#include "attention.cuh"
#include "attn_mma.cuh"
#include "tc_ptx.cuh"

namespace amc {
namespace {
namespace ap = attn_ptx;

struct Tc5Geom {
  int T, Tk, NT, rem, Tmain;      // tokens, keys padded to 16, 128-row main tiles per unit, leftover rows, main rows
  int h, d, units;
  int kbox_rows, kbox_n;          // K / V tile = kbox_n TMA boxes of kbox_rows rows
  int q_bytes, qlo_bytes, kv_bytes, stage_bytes;
  int oc, tmem_cols;              // O column offset inside the S region, TMEM allocation (power of two)
  int nch;                        // 32-column chunks per score row
  float sl2;                      // log2(e) / sqrt(dh)
};

constexpr int TC5_HDR = 1024;     // mbarriers + TMEM slot

// ---- PTX not in tc_ptx.cuh ------------------------------------------------------------------------------------
// shared-memory matrix descriptor with an explicit swizzle mode (2 = 128B, 4 = 64B, 6 = 32B)
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
template <int KD> __device__ __forceinline__ constexpr uint32_t sw_layout() { return KD == 1 ? 6u : (KD == 2 ? 4u : 2u); }
// kind::f16 instruction descriptor: fp32 accumulate, bf16 operands, separate operand majors (1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc2(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait5() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// ===============================================================================================================
// Forward
// ===============================================================================================================
struct FwdBars {
  uint64_t full[2], empty[2];     // input ring (TMA -> MMA / leftover warp ; MMA commit + leftover warp -> TMA)
  uint64_t s_full, p_full, o_full, t_empty, o_staged;
  uint32_t tmem_slot;
};

template <int KD, int RM>
__global__ void __launch_bounds__(RM + 64, (RM == 64) ? 4 : 2)
attn_tc5_fwd_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mQlo,
                    const __grid_constant__ CUtensorMap mKV, const __grid_constant__ CUtensorMap mO, const Tc5Geom gm,
                    bf16* __restrict__ out, float* __restrict__ lse) {
  constexpr int dh = 16 * KD, RB = 32 * KD, NSW = RM / 32;
  constexpr uint32_t LAY = sw_layout<KD>(), SBO = 8 * RB;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  FwdBars* bars = reinterpret_cast<FwdBars*>(smem);
  const uint32_t stage0 = smem_u32(smem + TC5_HDR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = gm.T, Tk = gm.Tk, NT = gm.NT;
  const bool has_lo = gm.rem > 0;
  auto q_tile = [&](int s) { return stage0 + (uint32_t)(s * gm.stage_bytes); };
  auto qlo_tile = [&](int s) { return q_tile(s) + (uint32_t)gm.q_bytes; };
  auto k_tile = [&](int s) { return qlo_tile(s) + (uint32_t)gm.qlo_bytes; };
  auto v_tile = [&](int s) { return k_tile(s) + (uint32_t)gm.kv_bytes; };

  if (tid == 0) {
    tma_prefetch_desc(&mQ); tma_prefetch_desc(&mQlo); tma_prefetch_desc(&mKV); tma_prefetch_desc(&mO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars->full + s, 1);
      mbar_init(bars->empty + s, has_lo ? 2 : 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, RM);
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->t_empty, NSW);
    mbar_init(&bars->o_staged, NSW);
    fence_barrier_init();
  }
  if (warp == NSW) tmem_alloc(&bars->tmem_slot, (uint32_t)gm.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == NSW) {
    // ============================ issuer: TMA loads, MMAs, TMA stores (one thread) ============================
    if (lane == 0) {
      auto issue_loads = [&](int u, int s) {
        const int b = u / gm.h, hh = u - b * gm.h;
        const int col = hh * dh;
        uint32_t bytes = (uint32_t)(NT * RM * RB + 2 * Tk * RB + (has_lo ? 16 * RB : 0));
        mbar_expect_tx(bars->full + s, bytes);
        for (int t = 0; t < NT; ++t)
          ap::tma_load_3d(&mQ, bars->full + s, q_tile(s) + (uint32_t)(t * 128 * RB), col, t * 128, b);
        if (has_lo) ap::tma_load_3d(&mQlo, bars->full + s, qlo_tile(s), col, gm.Tmain, b);
        for (int bx = 0; bx < gm.kbox_n; ++bx) {
          const uint32_t off = (uint32_t)(bx * gm.kbox_rows * RB);
          ap::tma_load_3d(&mKV, bars->full + s, k_tile(s) + off, gm.d + col, bx * gm.kbox_rows, b);
          ap::tma_load_3d(&mKV, bars->full + s, v_tile(s) + off, 2 * gm.d + col, bx * gm.kbox_rows, b);
        }
      };
      auto issue_store = [&](int u, int s) {
        const int b = u / gm.h, hh = u - b * gm.h;
        for (int t = 0; t < NT; ++t) ap::tma_store_3d(&mO, q_tile(s) + (uint32_t)(t * 128 * RB), hh * dh, t * 128, b);
        ap::bulk_commit();
      };
      constexpr uint32_t idPV = make_idesc2(128, dh, 0, 1);
      issue_loads(blockIdx.x, 0);
      int it = 0, u_prev = -1;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(bars->full + s, (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        for (int t = 0; t < NT; ++t) {
          const int n = it * NT + t;
          mbar_wait(&bars->t_empty, (uint32_t)((n & 1) ^ 1));       // the previous tile's O has left TMEM
          tc_fence_after();
          // S[128, Tk] = Q_tile K^T : K-major operands, one MMA per 16 head-dim columns and per <= 256 keys
          const uint32_t qa = q_tile(s) + (uint32_t)(t * 128 * RB), kb = k_tile(s);
          for (int nc = 0; nc < Tk; nc += 256) {
            const int N = min(256, Tk - nc);
            const uint32_t idS = make_idesc2(128, N, 0, 0);
#pragma unroll
            for (int ks = 0; ks < KD; ++ks)
              umma_bf16(tmem_base + (uint32_t)nc, make_desc_sw(qa + ks * 32, 16, SBO, LAY),
                        make_desc_sw(kb + (uint32_t)(nc * RB) + ks * 32, 16, SBO, LAY), idS, ks > 0 ? 1u : 0u);
          }
          umma_commit(&bars->s_full);
          if (t == 0) {
            // while the softmax runs: ship the previous unit's output, then refill its stage with the next unit
            if (u_prev >= 0) {
              mbar_wait(&bars->o_staged, (uint32_t)((it * NT - 1) & 1));
              issue_store(u_prev, s ^ 1);
              ap::bulk_wait_read0();
            }
            if (u + (int)gridDim.x < gm.units) {
              mbar_wait(bars->empty + (s ^ 1), (uint32_t)((((it + 1) >> 1) & 1) ^ 1));
              issue_loads(u + gridDim.x, s ^ 1);
            }
          }
          mbar_wait(&bars->p_full, (uint32_t)(n & 1));
          tc_fence_after();
          // O[128, dh] = P V : A = bf16 P in TMEM (8 columns per 16 keys), B = V tile read MN-major
          const uint32_t vb = v_tile(s);
          for (int j = 0; j < Tk / 16; ++j)
            umma_bf16_ts(tmem_base + (uint32_t)gm.oc, tmem_base + (uint32_t)(j * 8),
                         make_desc_sw(vb + (uint32_t)(j * 16 * RB), SBO, SBO, LAY), idPV, j > 0 ? 1u : 0u);
          umma_commit(&bars->o_full);
        }
        umma_commit(bars->empty + s);
        u_prev = u;
      }
      mbar_wait(&bars->o_staged, (uint32_t)((it * NT - 1) & 1));
      issue_store(u_prev, (it - 1) & 1);
      ap::bulk_wait_all0();
    }
  } else if (warp == NSW + 1) {
    // ============================ leftover rows: one mma.sync 16-row block per unit ============================
    if (has_lo) {
      constexpr int NBC = 3;
      const int g = lane >> 2, cb = (lane & 3) * 2;
      const int NQ = Tk / 16, last_k0 = ((NQ - 1) / NBC) * NBC;
      int it = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        const int s = it & 1;
        const int b = u / gm.h, hh = u - b * gm.h;
        mbar_wait(bars->full + s, (uint32_t)((it >> 1) & 1));
        const uint32_t qb = qlo_tile(s), kb = k_tile(s), vb = v_tile(s);
        uint32_t aq[KD][4];
#pragma unroll
        for (int ks = 0; ks < KD; ++ks) ap::ldsm_x4(aq[ks], ap::addrA<KD>(qb, 0, ks, lane));
        float o[2 * KD][4];
#pragma unroll
        for (int n = 0; n < 2 * KD; ++n) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int k0 = 0; k0 < last_k0; k0 += NBC)
          ap::fwd_chunk<KD, NBC, false>(kb, vb, k0, NBC, T, aq, o, m0, m1, l0, l1, gm.sl2, lane);
        ap::fwd_chunk<KD, NBC, true>(kb, vb, last_k0, NQ - last_k0, T, aq, o, m0, m1, l0, l1, gm.sl2, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(bars->empty + s);       // this warp is done with the stage's tiles
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        const int r0 = gm.Tmain + g, r1 = r0 + 8;
        bf16* ob = out + ((size_t)b * T) * gm.d + hh * dh + cb;
#pragma unroll
        for (int n = 0; n < 2 * KD; ++n) {
          if (r0 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * gm.d + n * 8) = ap::pack2(o[n][0] * i0, o[n][1] * i0);
          if (r1 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * gm.d + n * 8) = ap::pack2(o[n][2] * i1, o[n][3] * i1);
        }
        if (lse != nullptr && (lane & 3) == 0) {
          float* lp = lse + ((size_t)b * gm.h + hh) * T;
          if (r0 < T) lp[r0] = fmaf(m0, gm.sl2, __log2f(l0));
          if (r1 < T) lp[r1] = fmaf(m1, gm.sl2, __log2f(l1));
        }
      }
    }
  } else {
    // ============================ softmax warps: thread = query row = TMEM lane ============================
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sl2 = gm.sl2;
    const int nch = gm.nch;
    int it = 0;
    for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
      const int s = it & 1;
      const int b = u / gm.h, hh = u - b * gm.h;
      for (int t = 0; t < NT; ++t) {
        const int n = it * NT + t;
        mbar_wait(&bars->s_full, (uint32_t)(n & 1));
        tc_fence_after();
        uint32_t ra[32], rb[32];
        // ---- pass 1: row maximum (columns >= T are padding) ----
        float mx = -INFINITY;
        auto max32 = [&](uint32_t (&r)[32], int c) {
          if (c * 32 + 32 > T) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j >= T) r[j] = 0xff800000u;
          }
          float a0 = __uint_as_float(r[0]), a1 = __uint_as_float(r[1]), a2 = __uint_as_float(r[2]), a3 = __uint_as_float(r[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            a0 = fmaxf(a0, __uint_as_float(r[j])); a1 = fmaxf(a1, __uint_as_float(r[j + 1]));
            a2 = fmaxf(a2, __uint_as_float(r[j + 2])); a3 = fmaxf(a3, __uint_as_float(r[j + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
        };
        tmem_ld32(trow, ra);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait32(ra);
          if (c + 1 < nch) tmem_ld32(trow + (uint32_t)((c + 1) * 32), rb);
          max32(ra, c);
          if (c + 1 < nch) {
            tmem_ld_wait32(rb);
            if (c + 2 < nch) tmem_ld32(trow + (uint32_t)((c + 2) * 32), ra);
            max32(rb, c + 1);
          }
        }
        // ---- pass 2: p = exp2(s c - m c), row sum, bf16 P back into TMEM over the S columns ----
        const float ms = mx * sl2;
        float sum0 = 0.f, sum1 = 0.f;
        auto exp32 = [&](uint32_t (&r)[32], int c) {
          if (c * 32 + 32 > T) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j >= T) r[j] = 0xff800000u;
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ap::ex2(fmaf(__uint_as_float(r[2 * j]), sl2, -ms));
            const float p1 = ap::ex2(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -ms));
            sum0 += p0; sum1 += p1;
            pk[j] = ap::pack2(p0, p1);
          }
          tmem_st16(trow + (uint32_t)(c * 16), pk);
        };
        tmem_ld32(trow, ra);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait32(ra);
          if (c + 1 < nch) tmem_ld32(trow + (uint32_t)((c + 1) * 32), rb);
          exp32(ra, c);
          if (c + 1 < nch) {
            tmem_ld_wait32(rb);
            if (c + 2 < nch) tmem_ld32(trow + (uint32_t)((c + 2) * 32), ra);
            exp32(rb, c + 1);
          }
        }
        tmem_st_wait5();
        tc_fence_before();
        mbar_arrive(&bars->p_full);
        const float sum = sum0 + sum1, inv = 1.f / sum;
        const int rg = t * 128 + row;                         // row inside the frame
        if (lse != nullptr && rg < T) lse[((size_t)b * gm.h + hh) * T + rg] = ms + __log2f(sum);
        // ---- epilogue: O / sum -> bf16 -> staging (the dead Q tile) -> TMA store by the issuer ----
        mbar_wait(&bars->o_full, (uint32_t)(n & 1));
        tc_fence_after();
        const uint32_t ot = q_tile(s) + (uint32_t)(t * 128 * RB);
        if (KD == 1) {
          uint32_t o16[16];
          tmem_ld16(trow + (uint32_t)gm.oc, o16);
          tmem_ld_wait16(o16);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->t_empty);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch)
            sts128(ap::chunk_addr<KD>(ot, row, ch),
                   ap::pack2(__uint_as_float(o16[8 * ch]) * inv, __uint_as_float(o16[8 * ch + 1]) * inv),
                   ap::pack2(__uint_as_float(o16[8 * ch + 2]) * inv, __uint_as_float(o16[8 * ch + 3]) * inv),
                   ap::pack2(__uint_as_float(o16[8 * ch + 4]) * inv, __uint_as_float(o16[8 * ch + 5]) * inv),
                   ap::pack2(__uint_as_float(o16[8 * ch + 6]) * inv, __uint_as_float(o16[8 * ch + 7]) * inv));
        } else {
          tmem_ld32(trow + (uint32_t)gm.oc, ra);
          if (KD == 4) tmem_ld32(trow + (uint32_t)(gm.oc + 32), rb);
          tmem_ld_wait32(ra);
          if (KD == 4) tmem_ld_wait32(rb);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->t_empty);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            sts128(ap::chunk_addr<KD>(ot, row, ch),
                   ap::pack2(__uint_as_float(ra[8 * ch]) * inv, __uint_as_float(ra[8 * ch + 1]) * inv),
                   ap::pack2(__uint_as_float(ra[8 * ch + 2]) * inv, __uint_as_float(ra[8 * ch + 3]) * inv),
                   ap::pack2(__uint_as_float(ra[8 * ch + 4]) * inv, __uint_as_float(ra[8 * ch + 5]) * inv),
                   ap::pack2(__uint_as_float(ra[8 * ch + 6]) * inv, __uint_as_float(ra[8 * ch + 7]) * inv));
          if (KD == 4) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              sts128(ap::chunk_addr<KD>(ot, row, 4 + ch),
                     ap::pack2(__uint_as_float(rb[8 * ch]) * inv, __uint_as_float(rb[8 * ch + 1]) * inv),
                     ap::pack2(__uint_as_float(rb[8 * ch + 2]) * inv, __uint_as_float(rb[8 * ch + 3]) * inv),
                     ap::pack2(__uint_as_float(rb[8 * ch + 4]) * inv, __uint_as_float(rb[8 * ch + 5]) * inv),
                     ap::pack2(__uint_as_float(rb[8 * ch + 6]) * inv, __uint_as_float(rb[8 * ch + 7]) * inv));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_staged);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NSW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)gm.tmem_cols);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
constexpr size_t TC5_SMEM_MAX = 227 * 1024;

inline int up1024(int v) { return (v + 1023) / 1024 * 1024; }

// RM = main-tile rows (64 | 128); false = shape outside this kernel's envelope
bool tc5_plan(int B, int T, int h, int dh, Tc5Geom& g, int& RM) {
  if (!(dh == 16 || dh == 32 || dh == 64) || T < 49 || T > 272 || h < 1) return false;
  const int RB = 2 * dh;
  g.T = T; g.Tk = (T + 15) / 16 * 16; g.h = h; g.d = h * dh; g.units = B * h;
  if (T <= 80) { RM = 64; g.NT = 1; g.Tmain = std::min(T, 64); }
  else if (T <= 144) { RM = 128; g.NT = 1; g.Tmain = std::min(T, 128); }
  else { RM = 128; g.NT = 2; g.Tmain = std::min(T, 256); }
  g.rem = T - g.Tmain;
  g.kbox_n = g.Tk > 256 ? 2 : 1;
  g.kbox_rows = g.Tk / g.kbox_n;
  g.q_bytes = g.NT * 128 * RB;
  g.qlo_bytes = g.rem > 0 ? up1024(16 * RB) : 0;
  g.kv_bytes = up1024(g.Tk * RB);
  g.stage_bytes = g.q_bytes + g.qlo_bytes + 2 * g.kv_bytes;
  g.nch = (g.Tk + 31) / 32;
  g.oc = 16 * g.nch;
  const int need = std::max(32 * g.nch, g.oc + dh);
  g.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
  g.sl2 = 1.4426950408889634f / sqrtf((float)dh);
  return true;
}
size_t tc5_fwd_bytes(const Tc5Geom& g) { return 1024 + TC5_HDR + (size_t)2 * g.stage_bytes; }

}  // namespace

bool attn_tc5_supported(int T, int h, int dh) {
  Tc5Geom g;
  int RM;
  return tc5_plan(1, T, h, dh, g, RM) && tc5_fwd_bytes(g) <= TC5_SMEM_MAX;
}

int attn_tc5_fwd(int B, int T, int h, int dh, const bf16* qkv, bf16* out, float* lse, bool* handled, cudaStream_t st) {
  *handled = false;
  Tc5Geom g;
  int RM = 0;
  if (!tc5_plan(B, T, h, dh, g, RM)) return 0;
  const size_t sm = tc5_fwd_bytes(g);
  if (sm > TC5_SMEM_MAX || (g.d * 2) % 16 != 0) return 0;
  CUtensorMap mQ, mQlo, mKV, mO;
  AMC_TRY(attn_make_map3(&mQ, qkv, B, T, 3 * g.d, dh, RM));
  AMC_TRY(attn_make_map3(&mQlo, qkv, B, T, 3 * g.d, dh, 16));
  AMC_TRY(attn_make_map3(&mKV, qkv, B, T, 3 * g.d, dh, g.kbox_rows));
  AMC_TRY(attn_make_map3(&mO, out, B, T, g.d, dh, RM));
#define AMC_TC5_FWD(KD, RM_)                                                                                          \
  do {                                                                                                                \
    auto kern = attn_tc5_fwd_kernel<KD, RM_>;                                                                          \
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC5_SMEM_MAX));             \
    int occ = 1;                                                                                                      \
    AMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RM_ + 64, sm));                                \
    occ = std::max(1, std::min(occ, 512 / g.tmem_cols));                                                              \
    const int grid = std::min(g.units, attn_sm_count() * occ);                                                        \
    kern<<<grid, RM_ + 64, sm, st>>>(mQ, mQlo, mKV, mO, g, out, lse);                                                 \
  } while (0)
#define AMC_TC5_FWD_KD(KD)                  \
  do {                                      \
    if (RM == 64) AMC_TC5_FWD(KD, 64);      \
    else AMC_TC5_FWD(KD, 128);              \
  } while (0)
  if (dh == 16) AMC_TC5_FWD_KD(1);
  else if (dh == 32) AMC_TC5_FWD_KD(2);
  else AMC_TC5_FWD_KD(4);
#undef AMC_TC5_FWD_KD
#undef AMC_TC5_FWD
  AMC_LAUNCH_CHECK();
  *handled = true;
  return 0;
}

}  // namespace amc
