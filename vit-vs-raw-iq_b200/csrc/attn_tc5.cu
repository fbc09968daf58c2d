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
#include <map>
#include <mutex>

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
  int oc, tmem_cols, o_sep;       // O column offset, TMEM allocation (power of two), 1 = O outside the S columns
  int nst;                        // input stages (2..4)
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
constexpr int TC5_MAXST = 4;
struct FwdBars {
  uint64_t full[TC5_MAXST], empty[TC5_MAXST];   // input ring: TMA -> MMA / leftover warp ; MMA commit (+ leftover warp) -> TMA
  uint64_t o_staged[TC5_MAXST];                 // the unit's output rows are staged in its (dead) Q tile -> TMA store
  uint64_t s_full, p_full, o_full, t_empty;
  uint32_t tmem_slot;
};

// kernel-study trace (AMC_TC5_TRACE=1): clock64 stamps of CTA 0, [role][unit][event]
constexpr int TR_UNITS = 12, TR_EV = 8;
__device__ __forceinline__ void tr(long long* trace, int role, int it, int ev) {
  if (trace != nullptr && blockIdx.x == 0 && it < TR_UNITS) trace[(role * TR_UNITS + it) * TR_EV + ev] = clock64();
}

// threads: RM/32 softmax warps | MMA issuer warp | leftover warp | TMA warp
template <int KD, int RM>
__global__ void __launch_bounds__(RM + 96, (RM == 64) ? (KD == 4 ? 3 : 4) : 2)
attn_tc5_fwd_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mQlo,
                    const __grid_constant__ CUtensorMap mKV, const __grid_constant__ CUtensorMap mO, const Tc5Geom gm,
                    bf16* __restrict__ out, float* __restrict__ lse, long long* __restrict__ trace) {
  constexpr int dh = 16 * KD, RB = 32 * KD, NSW = RM / 32;
  constexpr uint32_t LAY = sw_layout<KD>(), SBO = 8 * RB;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  FwdBars* bars = reinterpret_cast<FwdBars*>(smem);
  const uint32_t stage0 = smem_u32(smem + TC5_HDR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = gm.T, Tk = gm.Tk, NT = gm.NT, NST = gm.nst;
  const bool has_lo = gm.rem > 0;
  auto q_tile = [&](int s) { return stage0 + (uint32_t)(s * gm.stage_bytes); };
  auto qlo_tile = [&](int s) { return q_tile(s) + (uint32_t)gm.q_bytes; };
  auto k_tile = [&](int s) { return qlo_tile(s) + (uint32_t)gm.qlo_bytes; };
  auto v_tile = [&](int s) { return k_tile(s) + (uint32_t)gm.kv_bytes; };

  if (tid == 0) {
    tma_prefetch_desc(&mQ); tma_prefetch_desc(&mQlo); tma_prefetch_desc(&mKV); tma_prefetch_desc(&mO);
    for (int s = 0; s < NST; ++s) {
      mbar_init(bars->full + s, 1);
      mbar_init(bars->empty + s, has_lo ? 2 : 1);
      mbar_init(bars->o_staged + s, NSW * NT);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, RM);
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->t_empty, NSW);
    fence_barrier_init();
  }
  if (warp == NSW) tmem_alloc(&bars->tmem_slot, (uint32_t)gm.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == NSW) {
    // ============================ MMA issuer (one thread) ============================
    if (lane == 0) {
      // descriptors differ only in the 14-bit start-address field: build the constant halves once
      constexpr uint32_t idPV = make_idesc2(128, dh, 0, 1);
      const uint64_t hiK = make_desc_sw(0, 16, SBO, LAY), hiV = make_desc_sw(0, SBO, SBO, LAY);
      const uint32_t idS0 = make_idesc2(128, min(256, Tk), 0, 0), idS1 = make_idesc2(128, max(16, Tk - 256), 0, 0);
      const int npv = Tk / 16;
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        mbar_wait(bars->full + s, ph);
        tc_fence_after();
        tr(trace, 0, it, 0);
        const uint32_t kb = k_tile(s) >> 4, vb = v_tile(s) >> 4;
        for (int t = 0; t < NT; ++t) {
          const int n = it * NT + t;
          if (!gm.o_sep) {                                           // O aliases the S columns: wait until it has been read
            mbar_wait(&bars->t_empty, (uint32_t)((n & 1) ^ 1));
            tc_fence_after();
          }
          // S[128, Tk] = Q_tile K^T : K-major operands, one MMA per 16 head-dim columns and per <= 256 keys
          const uint32_t qa = (q_tile(s) + (uint32_t)(t * 128 * RB)) >> 4;
#pragma unroll
          for (int ks = 0; ks < KD; ++ks)
            umma_bf16(tmem_base, hiK | (uint64_t)(qa + 2 * ks), hiK | (uint64_t)(kb + 2 * ks), idS0, ks > 0 ? 1u : 0u);
          if (Tk > 256) {
#pragma unroll
            for (int ks = 0; ks < KD; ++ks)
              umma_bf16(tmem_base + 256u, hiK | (uint64_t)(qa + 2 * ks), hiK | (uint64_t)(kb + 16 * RB + 2 * ks), idS1,
                        ks > 0 ? 1u : 0u);
          }
          umma_commit(&bars->s_full);
          if (t == 0) tr(trace, 0, it, 1);
          mbar_wait(&bars->p_full, (uint32_t)(n & 1));
          tc_fence_after();
          if (t == 0) tr(trace, 0, it, 2);
          // O[128, dh] = P V : A = bf16 P in TMEM (8 columns per 16 keys), B = V tile read MN-major (16 keys = RB * 16 bytes)
          const uint32_t to = tmem_base + (uint32_t)gm.oc;
          umma_bf16_ts(to, tmem_base, hiV | (uint64_t)vb, idPV, 0u);
#pragma unroll 4
          for (int j = 1; j < npv; ++j)
            umma_bf16_ts(to, tmem_base + (uint32_t)(j * 8), hiV | (uint64_t)(vb + j * RB), idPV, 1u);
          umma_commit(&bars->o_full);
        }
        umma_commit(bars->empty + s);
        tr(trace, 0, it, 3);
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == NSW + 2) {
    // ============================ TMA: loads run NST - 1 units ahead, stores trail the epilogue ============================
    if (lane == 0) {
      auto issue_loads = [&](int u, int s) {
        const int b = u / gm.h, hh = u - b * gm.h;
        const int col = hh * dh;
        mbar_expect_tx(bars->full + s, (uint32_t)(NT * RM * RB + 2 * Tk * RB + (has_lo ? 16 * RB : 0)));
        for (int t = 0; t < NT; ++t)
          ap::tma_load_3d(&mQ, bars->full + s, q_tile(s) + (uint32_t)(t * 128 * RB), col, t * 128, b);
        if (has_lo) ap::tma_load_3d(&mQlo, bars->full + s, qlo_tile(s), col, gm.Tmain, b);
        for (int bx = 0; bx < gm.kbox_n; ++bx) {
          const uint32_t off = (uint32_t)(bx * gm.kbox_rows * RB);
          ap::tma_load_3d(&mKV, bars->full + s, k_tile(s) + off, gm.d + col, bx * gm.kbox_rows, b);
          ap::tma_load_3d(&mKV, bars->full + s, v_tile(s) + off, 2 * gm.d + col, bx * gm.kbox_rows, b);
        }
      };
      for (int k = 0; k < NST; ++k)
        if ((int)blockIdx.x + k * (int)gridDim.x < gm.units) issue_loads(blockIdx.x + k * gridDim.x, k);
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        const int b = u / gm.h, hh = u - b * gm.h;
        mbar_wait(bars->o_staged + s, ph);                    // every softmax warp has staged its rows of this unit
        tr(trace, 1, it, 0);
        for (int t = 0; t < NT; ++t) ap::tma_store_3d(&mO, q_tile(s) + (uint32_t)(t * 128 * RB), hh * dh, t * 128, b);
        ap::bulk_commit();
        const int un = u + NST * (int)gridDim.x;
        if (un < gm.units) {
          ap::bulk_wait_read0();                              // the store has read the staging rows
          tr(trace, 1, it, 1);
          mbar_wait(bars->empty + s, ph);                     // MMAs (and the leftover warp) are done with K / V
          tr(trace, 1, it, 2);
          issue_loads(un, s);
        }
        if (++s == NST) { s = 0; ph ^= 1; }
      }
      ap::bulk_wait_all0();
    }
  } else if (warp == NSW + 1) {
    // ============================ leftover rows: one mma.sync 16-row block per unit ============================
    if (has_lo) {
      constexpr int NBC = 3;
      const int g = lane >> 2, cb = (lane & 3) * 2;
      const int NQ = Tk / 16, last_k0 = ((NQ - 1) / NBC) * NBC;
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        const int b = u / gm.h, hh = u - b * gm.h;
        mbar_wait(bars->full + s, ph);
        if (lane == 0) tr(trace, 3, it, 0);
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
        if (lane == 0) tr(trace, 3, it, 1);
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
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ============================ softmax warps: thread = query row = TMEM lane ============================
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sl2 = gm.sl2;
    const int nch = gm.nch;
    int it = 0, s = 0;
    for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
      const int b = u / gm.h, hh = u - b * gm.h;
      for (int t = 0; t < NT; ++t) {
        const int n = it * NT + t;
        mbar_wait(&bars->s_full, (uint32_t)(n & 1));
        tc_fence_after();
        if (tid == 0 && t == 0) tr(trace, 2, it, 0);
        uint32_t ra[32], rb[32];
        // ---- pass 1: row maximum (columns >= T are padding) ----
        float mx = -INFINITY;
        auto max32 = [&](uint32_t (&r)[32], int c) {
          const int tv = T - c * 32;                           // valid columns of this chunk (warp-uniform)
          if (tv < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j >= tv) r[j] = 0xff800000u;
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
        if (tid == 0 && t == 0) tr(trace, 2, it, 1);
        // ---- pass 2: p = exp2(s c - m c), row sum, bf16 P back into TMEM over the S columns ----
        const float ms = mx * sl2;
        float sum0 = 0.f, sum1 = 0.f;
        auto exp32 = [&](uint32_t (&r)[32], int c) {
          const int tv = T - c * 32;
          if (tv < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j >= tv) r[j] = 0xff800000u;
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
        if (tid == 0 && t == 0) tr(trace, 2, it, 2);
        const float sum = sum0 + sum1, inv = 1.f / sum;
        const int rg = t * 128 + row;                         // row inside the frame
        if (lse != nullptr && rg < T) lse[((size_t)b * gm.h + hh) * T + rg] = ms + __log2f(sum);
        // ---- epilogue: O / sum -> bf16 -> staging (the dead Q tile) -> TMA store by the TMA warp ----
        mbar_wait(&bars->o_full, (uint32_t)(n & 1));
        tc_fence_after();
        if (tid == 0 && t == 0) tr(trace, 2, it, 3);
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
        if (lane == 0) mbar_arrive(bars->o_staged + s);
        if (tid == 0 && t == 0) tr(trace, 2, it, 4);
      }
      if (++s == NST) s = 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NSW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)gm.tmem_cols);
  }
}

// ===============================================================================================================
// Backward (49 <= T <= 144: one 128-row query tile per unit + up to 16 leftover query rows)
// ===============================================================================================================
// Queries are the TMEM lanes.  Keys are walked in chunks of 64; per chunk
//   MMA thread   : S_c = Q K_c^T and dP_c = dO V_c^T                       (M = 128, N = chunk, K = dh) -> TMEM
//   row threads  : P = exp2(S c - lse2), dS = P * (dP - delta) / sqrt(dh)  -> bf16 tiles [query][key] in shared memory
//   MMA thread   : dV_c = P_c^T dO, dK_c = dS_c^T Q   (M = 64 keys, K = all queries: the tiles read MN-major)
//                  dQ  += dS_c K_c                    (M = 128, K = chunk: the same dS tile read K-major)
//   row threads  : dV_c / dK_c rows -> global ; after the last chunk dQ rows -> global
// delta = rowsum(dO * O) is taken once per unit by the row's thread (dO from the staged tile, O from global memory).
// The leftover query rows (T = 128 + 1) run on one mma.sync warp that writes its P / dS rows into rows 128.. of the
// same tiles, so the tensor-core dV / dK sums cover them, and keeps its own dQ rows in registers.
// TMEM: S_c [0, 64) and dP_c [64, 128), reused by dV_c [0, dh) and dK_c [64, 64 + dh) once the chunk has been read;
// dQ at [128, 128 + dh): 256 columns, two CTAs per SM.
constexpr int CK = 64;                 // keys per chunk
struct BwdBars {
  uint64_t full[TC5_MAXST], empty[TC5_MAXST];
  uint64_t sdp_full, ps_full, g_full, acc_free;
  uint32_t tmem_slot;
};
struct Tc5BwdGeom {
  int T, Tk, rem, h, d, units;
  int kbox_rows, kbox_n;
  int tile_bytes, stage_bytes, ps_bytes, nst, NC;
  float scale, sl2;
};

template <int KD>
__global__ void __launch_bounds__(224, 2)
attn_tc5_bwd_kernel(const __grid_constant__ CUtensorMap mQKV, const __grid_constant__ CUtensorMap mDO, const Tc5BwdGeom gm,
                    const bf16* __restrict__ out, const float* __restrict__ lse, bf16* __restrict__ dqkv,
                    long long* __restrict__ trace) {
  constexpr int dh = 16 * KD, RB = 32 * KD, NSW = 4;
  constexpr uint32_t LAY = sw_layout<KD>(), SBO = 8 * RB;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem);
  const uint32_t p_tile = smem_u32(smem + TC5_HDR), ds_tile = p_tile + (uint32_t)gm.ps_bytes;
  const uint32_t stage0 = ds_tile + (uint32_t)gm.ps_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = gm.T, Tk = gm.Tk, NST = gm.nst, NC = gm.NC;
  const bool has_lo = gm.rem > 0;
  auto q_tile = [&](int s) { return stage0 + (uint32_t)(s * gm.stage_bytes); };
  auto k_tile = [&](int s) { return q_tile(s) + (uint32_t)gm.tile_bytes; };
  auto v_tile = [&](int s) { return q_tile(s) + 2u * (uint32_t)gm.tile_bytes; };
  auto do_tile = [&](int s) { return q_tile(s) + 3u * (uint32_t)gm.tile_bytes; };

  if (tid == 0) {
    tma_prefetch_desc(&mQKV); tma_prefetch_desc(&mDO);
    for (int s = 0; s < NST; ++s) {
      mbar_init(bars->full + s, 1);
      mbar_init(bars->empty + s, has_lo ? 2 : 1);
    }
    mbar_init(&bars->sdp_full, 1);
    mbar_init(&bars->ps_full, NSW + (has_lo ? 1 : 0));
    mbar_init(&bars->g_full, 1);
    mbar_init(&bars->acc_free, NSW);
    fence_barrier_init();
  }
  if (warp == NSW) tmem_alloc(&bars->tmem_slot, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == NSW) {
    // ============================ MMA issuer (one thread) ============================
    if (lane == 0) {
      const uint64_t hiK = make_desc_sw(0, 16, SBO, LAY);          // K-major view of an input tile (rows of RB bytes)
      const uint64_t hiM = make_desc_sw(0, SBO, SBO, LAY);         // MN-major view of an input tile
      const uint64_t hiPK = make_desc_sw(0, 16, 1024, 2u);         // P / dS tile (128-byte rows), K-major
      const uint64_t hiPM = make_desc_sw(0, 1024, 1024, 2u);       // P / dS tile, MN-major (keys contiguous)
      constexpr uint32_t idG = make_idesc2(64, dh, 1, 1);          // dV_c / dK_c : both operands MN-major
      constexpr uint32_t idQ = make_idesc2(128, dh, 0, 1);         // dQ : A = dS K-major, B = K MN-major
      const int nq = Tk / 16;                                       // 16-row query steps of the dV / dK sums
      int it = 0, s = 0, cc = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        mbar_wait(bars->full + s, ph);
        tc_fence_after();
        tr(trace, 0, it, 0);
        const uint32_t qa = q_tile(s) >> 4, ka = k_tile(s) >> 4, va = v_tile(s) >> 4, da = do_tile(s) >> 4;
        for (int c = 0; c < NC; ++c, ++cc) {
          const int k0 = c * CK, wc = min(CK, Tk - k0);
          const uint32_t idS = make_idesc2(128, wc, 0, 0);
          mbar_wait(&bars->acc_free, (uint32_t)((cc & 1) ^ 1));     // the previous chunk's dV / dK (and dQ) have been read
          tc_fence_after();
          const uint32_t kr = (uint32_t)(k0 * RB) >> 4;
#pragma unroll
          for (int ks = 0; ks < KD; ++ks)
            umma_bf16(tmem_base, hiK | (uint64_t)(qa + 2 * ks), hiK | (uint64_t)(ka + kr + 2 * ks), idS, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < KD; ++ks)
            umma_bf16(tmem_base + 64u, hiK | (uint64_t)(da + 2 * ks), hiK | (uint64_t)(va + kr + 2 * ks), idS, ks > 0 ? 1u : 0u);
          umma_commit(&bars->sdp_full);
          if (c == 0) tr(trace, 0, it, 1);
          mbar_wait(&bars->ps_full, (uint32_t)(cc & 1));
          tc_fence_after();
          if (c == 0) tr(trace, 0, it, 2);
          // dV_c[64 keys, dh] = P_c^T dO ; dK_c = dS_c^T Q : 16 query rows per step (2048 B of the tiles, 16 * RB of dO / Q)
          const uint32_t pa = p_tile >> 4, sa = ds_tile >> 4;
#pragma unroll 3
          for (int j = 0; j < nq; ++j)
            umma_bf16(tmem_base, hiPM | (uint64_t)(pa + j * 128), hiM | (uint64_t)(da + j * RB), idG, j > 0 ? 1u : 0u);
#pragma unroll 3
          for (int j = 0; j < nq; ++j)
            umma_bf16(tmem_base + 64u, hiPM | (uint64_t)(sa + j * 128), hiM | (uint64_t)(qa + j * RB), idG, j > 0 ? 1u : 0u);
          // dQ[128, dh] += dS_c K_c : 16 keys per step (32 B inside the dS rows, 16 * RB of K)
          for (int ks = 0; ks < wc / 16; ++ks)
            umma_bf16(tmem_base + 128u, hiPK | (uint64_t)(sa + 2 * ks), hiM | (uint64_t)(ka + kr + ks * RB), idQ,
                      (c > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&bars->g_full);
        }
        umma_commit(bars->empty + s);
        tr(trace, 0, it, 3);
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == NSW + 2) {
    // ============================ TMA loads, NST units ahead ============================
    if (lane == 0) {
      auto issue_loads = [&](int u, int s) {
        const int b = u / gm.h, hh = u - b * gm.h;
        const int col = hh * dh;
        mbar_expect_tx(bars->full + s, (uint32_t)(4 * Tk * RB));
        for (int bx = 0; bx < gm.kbox_n; ++bx) {
          const uint32_t off = (uint32_t)(bx * gm.kbox_rows * RB);
          const int r0 = bx * gm.kbox_rows;
          ap::tma_load_3d(&mQKV, bars->full + s, q_tile(s) + off, col, r0, b);
          ap::tma_load_3d(&mQKV, bars->full + s, k_tile(s) + off, gm.d + col, r0, b);
          ap::tma_load_3d(&mQKV, bars->full + s, v_tile(s) + off, 2 * gm.d + col, r0, b);
          ap::tma_load_3d(&mDO, bars->full + s, do_tile(s) + off, col, r0, b);
        }
      };
      for (int k = 0; k < NST; ++k)
        if ((int)blockIdx.x + k * (int)gridDim.x < gm.units) issue_loads(blockIdx.x + k * gridDim.x, k);
      int s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x) {
        const int un = u + NST * (int)gridDim.x;
        if (un < gm.units) {
          mbar_wait(bars->empty + s, ph);
          issue_loads(un, s);
        }
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == NSW + 1) {
    // ============================ leftover query rows 128.. : one mma.sync 16-row block ============================
    if (has_lo) {
      const int g = lane >> 2, cb = (lane & 3) * 2;
      const int r0 = 128 + g, r1 = r0 + 8;
      int it = 0, s = 0, cc = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
        const int b = u / gm.h, hh = u - b * gm.h;
        // O fragments of rows r0 / r1 in the A-fragment layout (columns 16 ks + cb, +1 and + 8) for delta
        uint32_t ofr[KD][4];
        const bf16* op = out + ((size_t)b * T) * gm.d + hh * dh + cb;
#pragma unroll
        for (int ks = 0; ks < KD; ++ks) {
          ofr[ks][0] = r0 < T ? *reinterpret_cast<const uint32_t*>(op + (size_t)r0 * gm.d + 16 * ks) : 0u;
          ofr[ks][1] = r1 < T ? *reinterpret_cast<const uint32_t*>(op + (size_t)r1 * gm.d + 16 * ks) : 0u;
          ofr[ks][2] = r0 < T ? *reinterpret_cast<const uint32_t*>(op + (size_t)r0 * gm.d + 16 * ks + 8) : 0u;
          ofr[ks][3] = r1 < T ? *reinterpret_cast<const uint32_t*>(op + (size_t)r1 * gm.d + 16 * ks + 8) : 0u;
        }
        const float* lp = lse + ((size_t)b * gm.h + hh) * T;
        const float l0 = r0 < T ? lp[r0] : INFINITY, l1 = r1 < T ? lp[r1] : INFINITY;   // padded rows: P = 0
        mbar_wait(bars->full + s, ph);
        const uint32_t qb = q_tile(s), kb = k_tile(s), vb = v_tile(s), db = do_tile(s);
        uint32_t aq[KD][4], ad[KD][4];
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int ks = 0; ks < KD; ++ks) {
          ap::ldsm_x4(aq[ks], ap::addrA<KD>(qb, 128, ks, lane));
          ap::ldsm_x4(ad[ks], ap::addrA<KD>(db, 128, ks, lane));
          d0 += ap::bf_lo(ad[ks][0]) * ap::bf_lo(ofr[ks][0]) + ap::bf_hi(ad[ks][0]) * ap::bf_hi(ofr[ks][0]) +
                ap::bf_lo(ad[ks][2]) * ap::bf_lo(ofr[ks][2]) + ap::bf_hi(ad[ks][2]) * ap::bf_hi(ofr[ks][2]);
          d1 += ap::bf_lo(ad[ks][1]) * ap::bf_lo(ofr[ks][1]) + ap::bf_hi(ad[ks][1]) * ap::bf_hi(ofr[ks][1]) +
                ap::bf_lo(ad[ks][3]) * ap::bf_lo(ofr[ks][3]) + ap::bf_hi(ad[ks][3]) * ap::bf_hi(ofr[ks][3]);
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
        d0 *= gm.scale; d1 *= gm.scale;
        float dq[2 * KD][4];
#pragma unroll
        for (int n = 0; n < 2 * KD; ++n) { dq[n][0] = 0.f; dq[n][1] = 0.f; dq[n][2] = 0.f; dq[n][3] = 0.f; }
        for (int c = 0; c < NC; ++c, ++cc) {
          const int k0 = c * CK, wc = min(CK, Tk - k0);
          float st[8][4], dp[8][4];
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            st[n][0] = 0.f; st[n][1] = 0.f; st[n][2] = 0.f; st[n][3] = 0.f;
            dp[n][0] = 0.f; dp[n][1] = 0.f; dp[n][2] = 0.f; dp[n][3] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (16 * j < wc) {
#pragma unroll
              for (int ks = 0; ks < KD; ++ks) {
                uint32_t bfr[4];
                ap::ldsm_x4(bfr, ap::addrB<KD>(kb, k0 + 16 * j, ks, lane));
                ap::mma_bf16(st[2 * j], aq[ks], bfr[0], bfr[1]);
                ap::mma_bf16(st[2 * j + 1], aq[ks], bfr[2], bfr[3]);
                ap::ldsm_x4(bfr, ap::addrB<KD>(vb, k0 + 16 * j, ks, lane));
                ap::mma_bf16(dp[2 * j], ad[ks], bfr[0], bfr[1]);
                ap::mma_bf16(dp[2 * j + 1], ad[ks], bfr[2], bfr[3]);
              }
            }
          }
          // the P / dS tiles are free once the previous chunk's gradient MMAs have retired
          mbar_wait(&bars->g_full, (uint32_t)((cc & 1) ^ 1));
          uint32_t sa[8][2];
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float p0 = ap::ex2(fmaf(st[n][0], gm.sl2, -l0)), p1 = ap::ex2(fmaf(st[n][1], gm.sl2, -l0));
            const float p2 = ap::ex2(fmaf(st[n][2], gm.sl2, -l1)), p3 = ap::ex2(fmaf(st[n][3], gm.sl2, -l1));
            sa[n][0] = ap::pack2(p0 * fmaf(dp[n][0], gm.scale, -d0), p1 * fmaf(dp[n][1], gm.scale, -d0));
            sa[n][1] = ap::pack2(p2 * fmaf(dp[n][2], gm.scale, -d1), p3 * fmaf(dp[n][3], gm.scale, -d1));
            ap::sts32(ap::chunk_addr<4>(p_tile, r0, n) + cb * 2, ap::pack2(p0, p1));
            ap::sts32(ap::chunk_addr<4>(p_tile, r1, n) + cb * 2, ap::pack2(p2, p3));
            ap::sts32(ap::chunk_addr<4>(ds_tile, r0, n) + cb * 2, sa[n][0]);
            ap::sts32(ap::chunk_addr<4>(ds_tile, r1, n) + cb * 2, sa[n][1]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->ps_full);
          // dQ rows += dS K_c
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (16 * j < wc) {
              const uint32_t a4[4] = {sa[2 * j][0], sa[2 * j][1], sa[2 * j + 1][0], sa[2 * j + 1][1]};
#pragma unroll
              for (int np = 0; np < KD; ++np) {
                uint32_t bfr[4];
                ap::ldsm_x4_t(bfr, ap::addrA<KD>(kb, k0 + 16 * j, np, lane));
                ap::mma_bf16(dq[2 * np], a4, bfr[0], bfr[1]);
                ap::mma_bf16(dq[2 * np + 1], a4, bfr[2], bfr[3]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars->empty + s);
        bf16* gp = dqkv + ((size_t)b * T) * (3 * gm.d) + hh * dh + cb;
#pragma unroll
        for (int n = 0; n < 2 * KD; ++n) {
          if (r0 < T) *reinterpret_cast<uint32_t*>(gp + (size_t)r0 * (3 * gm.d) + n * 8) = ap::pack2(dq[n][0], dq[n][1]);
          if (r1 < T) *reinterpret_cast<uint32_t*>(gp + (size_t)r1 * (3 * gm.d) + n * 8) = ap::pack2(dq[n][2], dq[n][3]);
        }
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ============================ row threads: thread = query row = TMEM lane ============================
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sl2 = gm.sl2, scale = gm.scale;
    const uint32_t prow = p_tile + (uint32_t)(row * 128), srow = ds_tile + (uint32_t)(row * 128);
    const int sw = row & 7;
    int it = 0, s = 0, cc = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < gm.units; u += gridDim.x, ++it) {
      const int b = u / gm.h, hh = u - b * gm.h;
      // delta = rowsum(dO * O) / sqrt(dh), lse2 of this row (rows >= T: lse2 = +inf -> P = 0)
      uint4 ov[2 * KD];
      float l2 = INFINITY;
      if (row < T) {
        const uint4* op = reinterpret_cast<const uint4*>(out + ((size_t)b * T + row) * gm.d + hh * dh);
#pragma unroll
        for (int ch = 0; ch < 2 * KD; ++ch) ov[ch] = __ldg(op + ch);
        l2 = __ldg(lse + ((size_t)b * gm.h + hh) * T + row);
      } else {
#pragma unroll
        for (int ch = 0; ch < 2 * KD; ++ch) ov[ch] = make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(bars->full + s, ph);
      float dl = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2 * KD; ++ch) {
        const uint4 a = ap::lds128(ap::chunk_addr<KD>(do_tile(s), row, ch));
        const uint4 o4 = ov[ch];
        dl += ap::bf_lo(a.x) * ap::bf_lo(o4.x) + ap::bf_hi(a.x) * ap::bf_hi(o4.x) + ap::bf_lo(a.y) * ap::bf_lo(o4.y) +
              ap::bf_hi(a.y) * ap::bf_hi(o4.y) + ap::bf_lo(a.z) * ap::bf_lo(o4.z) + ap::bf_hi(a.z) * ap::bf_hi(o4.z) +
              ap::bf_lo(a.w) * ap::bf_lo(o4.w) + ap::bf_hi(a.w) * ap::bf_hi(o4.w);
      }
      const float dls = dl * scale;
      if (tid == 0) tr(trace, 2, it, 0);
      for (int c = 0; c < NC; ++c, ++cc) {
        const int k0 = c * CK, wc = min(CK, Tk - k0);
        mbar_wait(&bars->sdp_full, (uint32_t)(cc & 1));
        tc_fence_after();
        if (tid == 0 && c == 0) tr(trace, 2, it, 1);
        // the P / dS tiles are free: the previous chunk's gradient MMAs retired before its accumulators were read
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf * 32 < wc) {
            uint32_t rs[32], rd[32];
            tmem_ld32(trow + (uint32_t)(hf * 32), rs);
            tmem_ld32(trow + 64u + (uint32_t)(hf * 32), rd);
            tmem_ld_wait32(rs);
            tmem_ld_wait32(rd);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {                  // 8 keys = one 16-byte chunk of the tile rows
              uint32_t pw[4], dw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = q4 * 8 + e * 2;
                const float p0 = ap::ex2(fmaf(__uint_as_float(rs[j]), sl2, -l2));
                const float p1 = ap::ex2(fmaf(__uint_as_float(rs[j + 1]), sl2, -l2));
                pw[e] = ap::pack2(p0, p1);
                dw[e] = ap::pack2(p0 * fmaf(__uint_as_float(rd[j]), scale, -dls), p1 * fmaf(__uint_as_float(rd[j + 1]), scale, -dls));
              }
              const uint32_t off = (uint32_t)(((hf * 4 + q4) ^ sw) << 4);
              sts128(prow + off, pw[0], pw[1], pw[2], pw[3]);
              sts128(srow + off, dw[0], dw[1], dw[2], dw[3]);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->ps_full);
        if (tid == 0 && c == 0) tr(trace, 2, it, 2);
        // ---- dV_c / dK_c rows (M = 64 accumulators: key 16 * warp + lane lives in lane < 16 of this quadrant) ----
        mbar_wait(&bars->g_full, (uint32_t)(cc & 1));
        tc_fence_after();
        if (tid == 0 && c == 0) tr(trace, 2, it, 3);
        const int key = k0 + 16 * warp + lane;
        const bool kv_ok = lane < 16 && key < T;
        bf16* gk = dqkv + ((size_t)b * T + key) * (3 * gm.d) + gm.d + hh * dh;
        if (KD == 1) {
          uint32_t a[16], bq[16];
          tmem_ld16(trow, a);
          tmem_ld16(trow + 64u, bq);
          tmem_ld_wait16(a);
          tmem_ld_wait16(bq);
          if (kv_ok) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              *reinterpret_cast<uint4*>(gk + gm.d + ch * 8) =
                  make_uint4(ap::pack2(__uint_as_float(a[8 * ch]), __uint_as_float(a[8 * ch + 1])),
                             ap::pack2(__uint_as_float(a[8 * ch + 2]), __uint_as_float(a[8 * ch + 3])),
                             ap::pack2(__uint_as_float(a[8 * ch + 4]), __uint_as_float(a[8 * ch + 5])),
                             ap::pack2(__uint_as_float(a[8 * ch + 6]), __uint_as_float(a[8 * ch + 7])));
              *reinterpret_cast<uint4*>(gk + ch * 8) =
                  make_uint4(ap::pack2(__uint_as_float(bq[8 * ch]), __uint_as_float(bq[8 * ch + 1])),
                             ap::pack2(__uint_as_float(bq[8 * ch + 2]), __uint_as_float(bq[8 * ch + 3])),
                             ap::pack2(__uint_as_float(bq[8 * ch + 4]), __uint_as_float(bq[8 * ch + 5])),
                             ap::pack2(__uint_as_float(bq[8 * ch + 6]), __uint_as_float(bq[8 * ch + 7])));
            }
          }
        } else {
#pragma unroll
          for (int q2 = 0; q2 < KD / 2; ++q2) {
            uint32_t a[32], bq[32];
            tmem_ld32(trow + (uint32_t)(q2 * 32), a);
            tmem_ld32(trow + 64u + (uint32_t)(q2 * 32), bq);
            tmem_ld_wait32(a);
            tmem_ld_wait32(bq);
            if (kv_ok) {
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                *reinterpret_cast<uint4*>(gk + gm.d + q2 * 32 + ch * 8) =
                    make_uint4(ap::pack2(__uint_as_float(a[8 * ch]), __uint_as_float(a[8 * ch + 1])),
                               ap::pack2(__uint_as_float(a[8 * ch + 2]), __uint_as_float(a[8 * ch + 3])),
                               ap::pack2(__uint_as_float(a[8 * ch + 4]), __uint_as_float(a[8 * ch + 5])),
                               ap::pack2(__uint_as_float(a[8 * ch + 6]), __uint_as_float(a[8 * ch + 7])));
                *reinterpret_cast<uint4*>(gk + q2 * 32 + ch * 8) =
                    make_uint4(ap::pack2(__uint_as_float(bq[8 * ch]), __uint_as_float(bq[8 * ch + 1])),
                               ap::pack2(__uint_as_float(bq[8 * ch + 2]), __uint_as_float(bq[8 * ch + 3])),
                               ap::pack2(__uint_as_float(bq[8 * ch + 4]), __uint_as_float(bq[8 * ch + 5])),
                               ap::pack2(__uint_as_float(bq[8 * ch + 6]), __uint_as_float(bq[8 * ch + 7])));
              }
            }
          }
        }
        if (c == NC - 1) {
          // ---- dQ row ----
          bf16* gq = dqkv + ((size_t)b * T + row) * (3 * gm.d) + hh * dh;
          if (KD == 1) {
            uint32_t a[16];
            tmem_ld16(trow + 128u, a);
            tmem_ld_wait16(a);
            if (row < T) {
#pragma unroll
              for (int ch = 0; ch < 2; ++ch)
                *reinterpret_cast<uint4*>(gq + ch * 8) =
                    make_uint4(ap::pack2(__uint_as_float(a[8 * ch]), __uint_as_float(a[8 * ch + 1])),
                               ap::pack2(__uint_as_float(a[8 * ch + 2]), __uint_as_float(a[8 * ch + 3])),
                               ap::pack2(__uint_as_float(a[8 * ch + 4]), __uint_as_float(a[8 * ch + 5])),
                               ap::pack2(__uint_as_float(a[8 * ch + 6]), __uint_as_float(a[8 * ch + 7])));
            }
          } else {
#pragma unroll
            for (int q2 = 0; q2 < KD / 2; ++q2) {
              uint32_t a[32];
              tmem_ld32(trow + 128u + (uint32_t)(q2 * 32), a);
              tmem_ld_wait32(a);
              if (row < T) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                  *reinterpret_cast<uint4*>(gq + q2 * 32 + ch * 8) =
                      make_uint4(ap::pack2(__uint_as_float(a[8 * ch]), __uint_as_float(a[8 * ch + 1])),
                                 ap::pack2(__uint_as_float(a[8 * ch + 2]), __uint_as_float(a[8 * ch + 3])),
                                 ap::pack2(__uint_as_float(a[8 * ch + 4]), __uint_as_float(a[8 * ch + 5])),
                                 ap::pack2(__uint_as_float(a[8 * ch + 6]), __uint_as_float(a[8 * ch + 7])));
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_free);
        if (tid == 0 && c == 0) tr(trace, 2, it, 4);
      }
      if (++s == NST) { s = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NSW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
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
  // O outside the S columns (the next tile's S MMA need not wait for the epilogue) when that costs no larger allocation
  auto pow2 = [](int need) { return need <= 128 ? 128 : (need <= 256 ? 256 : 512); };
  const int alias_cols = pow2(std::max(32 * g.nch, 16 * g.nch + dh)), sep_cols = pow2(32 * g.nch + dh);
  g.o_sep = sep_cols == alias_cols ? 1 : 0;
  g.oc = g.o_sep ? 32 * g.nch : 16 * g.nch;
  g.tmem_cols = alias_cols;
  g.sl2 = 1.4426950408889634f / sqrtf((float)dh);
  // CTAs per SM allowed by TMEM (and the warp budget), then as many input stages as their shared-memory share holds
  int ctas = std::min(512 / g.tmem_cols, RM == 64 ? 4 : 2);
  for (;; --ctas) {
    const int budget = (228 * 1024 - ctas * 1024) / ctas - 1024 - TC5_HDR;
    g.nst = std::min(TC5_MAXST, budget / g.stage_bytes);
    if (g.nst >= 2 || ctas == 1) break;
  }
  return g.nst >= 2;
}
size_t tc5_fwd_bytes(const Tc5Geom& g) { return 1024 + TC5_HDR + (size_t)g.nst * g.stage_bytes; }


bool tc5_bwd_plan(int B, int T, int h, int dh, Tc5BwdGeom& g) {
  if (!(dh == 16 || dh == 32 || dh == 64) || T < 49 || T > 144 || h < 1) return false;
  const int RB = 2 * dh;
  g.T = T; g.Tk = (T + 15) / 16 * 16; g.h = h; g.d = h * dh; g.units = B * h;
  g.rem = std::max(0, T - 128);
  g.kbox_n = 1; g.kbox_rows = g.Tk;
  // every tile holds at least the 128 rows an M = 128 operand descriptor (and the 128 row threads) touch
  g.tile_bytes = up1024(std::max(g.Tk, 128) * RB);
  g.stage_bytes = 4 * g.tile_bytes;
  g.ps_bytes = up1024(std::max(g.Tk, 128) * 128);
  g.NC = (g.Tk + CK - 1) / CK;
  g.scale = 1.f / sqrtf((float)dh);
  g.sl2 = 1.4426950408889634f * g.scale;
  for (int ctas = 2; ctas >= 1; --ctas) {
    const int budget = (228 * 1024 - ctas * 1024) / ctas - 1024 - TC5_HDR - 2 * g.ps_bytes;
    g.nst = std::min(TC5_MAXST, budget / g.stage_bytes);
    if (g.nst >= 2) break;
  }
  return g.nst >= 2;
}
size_t tc5_bwd_bytes(const Tc5BwdGeom& g) { return 1024 + TC5_HDR + (size_t)2 * g.ps_bytes + (size_t)g.nst * g.stage_bytes; }

// CTAs of one kernel an SM can hold: TMEM columns, shared memory (1 KB reserved per CTA), registers (allocated per warp
// in units of 256, 64 K per SM and 16 K per sub-partition), 2048 threads.
int tc5_ctas_per_sm(const void* kern, int threads, size_t smem, int tmem_cols) {
  static std::mutex mu;
  static std::map<const void*, int> regs_of;
  int regs;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto f = regs_of.find(kern);
    if (f == regs_of.end()) {
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return 1;
      f = regs_of.emplace(kern, fa.numRegs).first;
    }
    regs = f->second;
  }
  const int warps = (threads + 31) / 32, regs_warp = (regs * 32 + 255) / 256 * 256;
  const int by_regs = (4 * (16384 / regs_warp)) / ((warps + 3) / 4 * 4);
  const int by_smem = (int)((228 * 1024) / (smem + 1024));
  return std::max(1, std::min(std::min(512 / tmem_cols, by_smem), std::min(by_regs, 2048 / threads)));
}

// ---- kernel-study trace: AMC_TC5_TRACE=1 prints the clock64 stamps of CTA 0 after every launch (debug only) ----
long long* tc5_trace_begin() {
  static const bool on = [] { const char* e = getenv("AMC_TC5_TRACE"); return e && e[0] == '1'; }();
  if (!on) return nullptr;
  static long long* buf = nullptr;
  const size_t n = 4 * TR_UNITS * TR_EV * sizeof(long long);
  if (!buf && cudaMalloc(&buf, n) != cudaSuccess) return nullptr;
  cudaMemset(buf, 0, n);
  return buf;
}
void tc5_trace_end(long long* buf, const char* what, int T, int dh, cudaStream_t st) {
  long long h[4 * TR_UNITS * TR_EV];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
  long long t0 = 0;
  for (size_t i = 0; i < sizeof(h) / sizeof(h[0]); ++i)
    if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
  static const char* roles[4] = {"mma ", "tma ", "smx0", "left"};
  fprintf(stderr, "[tc5 trace %s T=%d dh=%d] cycles since the first stamp, per unit of CTA 0\n", what, T, dh);
  for (int r = 0; r < 4; ++r)
    for (int u = 0; u < TR_UNITS; ++u) {
      bool any = false;
      for (int e = 0; e < TR_EV; ++e) any |= h[(r * TR_UNITS + u) * TR_EV + e] != 0;
      if (!any) continue;
      fprintf(stderr, "  %s u%-2d:", roles[r], u);
      for (int e = 0; e < TR_EV; ++e) {
        const long long v = h[(r * TR_UNITS + u) * TR_EV + e];
        if (v) fprintf(stderr, " e%d=%lld", e, v - t0);
      }
      fprintf(stderr, "\n");
    }
}

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
  long long* trace = tc5_trace_begin();
#define AMC_TC5_FWD(KD, RM_)                                                                                          \
  do {                                                                                                                \
    auto kern = attn_tc5_fwd_kernel<KD, RM_>;                                                                          \
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC5_SMEM_MAX));             \
    /* (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that allocates TMEM, whatever its    \
       resources; the hardware does co-schedule such CTAs -- measured -- so the limits are applied here) */           \
    int occ = tc5_ctas_per_sm((const void*)kern, RM_ + 96, sm, g.tmem_cols);                                          \
    if (getenv("AMC_TC5_OCC")) occ = atoi(getenv("AMC_TC5_OCC"));                                                     \
    const int grid = std::min(g.units, attn_sm_count() * occ);                                                        \
    if (trace) {                                                                                                      \
      cudaFuncAttributes fa;                                                                                          \
      cudaFuncGetAttributes(&fa, kern);                                                                               \
      fprintf(stderr, "[tc5 fwd] grid %d occ %d smem %zu nst %d tmem %d o_sep %d units %d | regs %d static smem %zu local %zu maxdyn %d\n", \
              grid, occ, sm, g.nst, g.tmem_cols, g.o_sep, g.units, fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes); \
    }                                                                                                                 \
    kern<<<grid, RM_ + 96, sm, st>>>(mQ, mQlo, mKV, mO, g, out, lse, trace);                                                 \
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
  if (trace) tc5_trace_end(trace, "fwd", T, dh, st);
  *handled = true;
  return 0;
}

int attn_tc5_bwd(int B, int T, int h, int dh, const bf16* qkv, const bf16* out, const float* lse, const bf16* dout,
                 bf16* dqkv, bool* handled, cudaStream_t st) {
  *handled = false;
  Tc5BwdGeom g;
  if (out == nullptr || lse == nullptr || !tc5_bwd_plan(B, T, h, dh, g)) return 0;
  const size_t sm = tc5_bwd_bytes(g);
  if (sm > TC5_SMEM_MAX || (g.d * 2) % 16 != 0) return 0;
  CUtensorMap mQKV, mDO;
  AMC_TRY(attn_make_map3(&mQKV, qkv, B, T, 3 * g.d, dh, g.kbox_rows));
  AMC_TRY(attn_make_map3(&mDO, dout, B, T, g.d, dh, g.kbox_rows));
  long long* trace = tc5_trace_begin();
#define AMC_TC5_BWD(KD)                                                                                               \
  do {                                                                                                                \
    auto kern = attn_tc5_bwd_kernel<KD>;                                                                               \
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC5_SMEM_MAX));             \
    int occ = tc5_ctas_per_sm((const void*)kern, 224, sm, 256);                                                       \
    if (getenv("AMC_TC5_OCC")) occ = atoi(getenv("AMC_TC5_OCC"));                                                     \
    const int grid = std::min(g.units, attn_sm_count() * occ);                                                        \
    if (trace) fprintf(stderr, "[tc5 bwd] grid %d occ %d smem %zu nst %d units %d\n", grid, occ, sm, g.nst, g.units); \
    kern<<<grid, 224, sm, st>>>(mQKV, mDO, g, out, lse, dqkv, trace);                                                 \
  } while (0)
  if (dh == 16) AMC_TC5_BWD(1);
  else if (dh == 32) AMC_TC5_BWD(2);
  else AMC_TC5_BWD(4);
#undef AMC_TC5_BWD
  AMC_LAUNCH_CHECK();
  if (trace) tc5_trace_end(trace, "bwd", T, dh, st);
  *handled = true;
  return 0;
}

}  // namespace amc
