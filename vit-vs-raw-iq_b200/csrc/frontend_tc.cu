// Fused embedding front end (bf16 path): dataset z-score + framing + patchify + embedding GEMM + bias +
// positional encoding + dropout in ONE kernel -- SURVEY §8 rows a1-a5:
//   R/dataloader/dataset.py:215-222, V/dataloader/dataset.py:211-224 (normalise, transpose / cat(I,Q).view(1,32,64))
//   R/models/embedding/patch_embedding.py:38-60, V/models/embedding/patch_embedding.py:9-15 (Conv1d / Conv2d as a GEMM)
//   R/models/encoder.py:104-111 (CLS row shift, + encoding[:T], dropout)
//
// The IQ samples are read from HBM exactly once, as coalesced 128-bit loads (every warp-wide load is 512 contiguous
// bytes of a frame), normalised, converted to bf16 and written by the converter warps straight into the
// 128B-swizzled K-major shared-memory tile that tcgen05.mma consumes as its A operand: the patchified operand never
// exists in HBM in inference (training also streams it out once, for the embedding weight gradient).  The embedding
// weight W[d, K] is TMA-loaded once per CTA and stays resident; accumulators are double-buffered in TMEM so the
// epilogue (bias + positional encoding + dropout, CLS-shifted fp32 and bf16 rows) of tile i overlaps tile i+1.
//
// Warp roles (448 threads): warp 0 = W loader (TMA), warp 1 = MMA issuer (one thread), warps 2-5 = converters,
// warps 6-13 = epilogue (two per TMEM lane quadrant).  Bound: HBM (8 B per IQ sample in, 6 B per embedding
// element out); the GEMM (K <= 256) is a small fraction of the tensor pipe.
#include <algorithm>

#include "gemm_common.cuh"
#include "rowops.cuh"
#include "tc_ptx.cuh"

namespace amc {
namespace {

constexpr int FBM = 128, FBK = 64;
constexpr int F_NCONV = 4, F_NEPI = 8;
constexpr int F_THREADS = 64 + 32 * (F_NCONV + F_NEPI);
constexpr int F_CONV_WARP0 = 2, F_EPI_WARP0 = 2 + F_NCONV;
constexpr int F_STG_LD = 32;

struct FrontParams {
  int kind, raw;          // AMC_KIND_*, input layout (1 = dataset [B,L,2] interleaved, 0 = model layout)
  int B, M;               // frames, token rows B * Ttok
  int Ttok, K, d;
  int kb_total;           // ceil(K / 64)
  int S, L;               // raw-IQ: segment size, samples per frame
  int p;                  // ViT: patch size (image is 32 x 64)
  float mean[2], inv_std[2];
  int tiles_m, tiles_n, vec_ok;
  bf16* Aout;             // nullable: the patchified operand [M, K] (training: embedding weight gradient)
};

template <int BN> struct FrontCfg {
  static constexpr int A_BYTES = FBM * FBK * 2;                    // 16 KB per k-block
  static constexpr int W_BLOCK = BN * FBK * 2;                     // one k-block of W
  static constexpr int STG_BYTES = F_NEPI * 32 * F_STG_LD * 4;     // 32 KB
  static constexpr int AUX_BYTES = 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static size_t smem(int kb_total, int nstage) {
    return 1024 + (size_t)kb_total * W_BLOCK + (size_t)nstage * A_BYTES + STG_BYTES + AUX_BYTES;
  }
};

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// byte offset of element (row, kcol) inside a 128-row x 64-column K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t a_off(int row, int kcol) {
  return (uint32_t)(row * 128 + ((((kcol >> 3) ^ (row & 7)) << 4) | ((kcol & 7) << 1)));
}

// One converter warp's share of k-block `kb` of the A tile starting at token row m0.  Every iteration is one
// warp-wide 512-byte load (lane = one float4) followed by the shared-memory scatter of its 4 values.
__device__ __forceinline__ void convert_block(const FrontParams& p, const float* __restrict__ src, uint32_t sa, int m0,
                                              int kb, int cw, int lane) {
  if (p.kind == AMC_KIND_RAWIQ) {
    const int S = p.S;
    if (p.raw) {
      // tile = 128 * S consecutive (I, Q) pairs of the frame stream; float4 = samples n, n + 1
      const int nwl = 2 * S;
      const float4* base = reinterpret_cast<const float4*>(src + (size_t)m0 * S * 2);
      for (int wl = cw; wl < nwl; wl += F_NCONV) {
        const int n = wl * 64 + 2 * lane, row = n / S, s = n - row * S;
        if (m0 + row < p.M) {
          const float4 v = __ldg(base + wl * 32 + lane);
          const uint32_t wi = pk((v.x - p.mean[0]) * p.inv_std[0], (v.z - p.mean[0]) * p.inv_std[0]);
          const uint32_t wq = pk((v.y - p.mean[1]) * p.inv_std[1], (v.w - p.mean[1]) * p.inv_std[1]);
          sts32(sa + a_off(row, s), wi);
          sts32(sa + a_off(row, S + s), wq);
          if (p.Aout) {
            bf16* ao = p.Aout + (size_t)(m0 + row) * p.K;
            *reinterpret_cast<uint32_t*>(ao + s) = wi;
            *reinterpret_cast<uint32_t*>(ao + S + s) = wq;
          }
        } else {
          sts32(sa + a_off(row, s), 0u);
          sts32(sa + a_off(row, S + s), 0u);
        }
      }
    } else {
      // model layout [B, 2, L] (already normalised): per channel 128 * S consecutive samples per frame segment
      const int nwl = 2 * S;     // S warp-loads (128 floats each) per channel
      for (int wl = cw; wl < nwl; wl += F_NCONV) {
        const int c = wl / S, n = (wl - c * S) * 128 + lane * 4, row = n / S, s = n - row * S;
        const int gr = m0 + row;
        uint32_t w0 = 0u, w1 = 0u;
        if (gr < p.M) {
          const int b = gr / p.Ttok, t = gr - b * p.Ttok;
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)b * 2 + c) * p.L + (size_t)t * S + s));
          w0 = pk(v.x, v.y);
          w1 = pk(v.z, v.w);
          if (p.Aout) *reinterpret_cast<uint2*>(p.Aout + (size_t)gr * p.K + c * S + s) = make_uint2(w0, w1);
        }
        sts64(sa + a_off(row, c * S + s), w0, w1);
      }
    }
  } else {
    // ViT, image 32 x 64, patch p in {4, 8, 16}: token = ph * (64/p) + pw, k = r * p + cc  (patch_embedding.py:12-14)
    const int pp = p.p, wp = 64 / pp;
    const int rpk = min(pp, 64 / pp);              // patch rows r per 64-wide k-block
    const int fpt = FBM / p.Ttok;                  // frames per tile
    const int f0 = m0 / p.Ttok;
    if (p.raw) {
      // dataset layout: image = cat(I, Q).view(32, 64) -> image rows 0-15 = I samples, 16-31 = Q samples
      // (V/dataloader/dataset.py:216-224); one warp-load = one image row of I and the matching row of Q
      const int nph = 16 / pp, nwl = fpt * nph * rpk;
      for (int wl = cw; wl < nwl; wl += F_NCONV) {
        const int rr = wl % rpk, t2 = wl / rpk, ph = t2 % nph, f = t2 / nph;
        const int r = kb * rpk + rr, x = 2 * lane, pw = x / pp, cc = x - pw * pp;
        const int rowI = f * p.Ttok + ph * wp + pw, rowQ = rowI + nph * wp, kcol = rr * pp + cc;
        uint32_t wi = 0u, wq = 0u;
        if (f0 + f < p.B) {
          const int n = (ph * pp + r) * 64 + x;
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(f0 + f) * 1024 + n) * 2));
          wi = pk((v.x - p.mean[0]) * p.inv_std[0], (v.z - p.mean[0]) * p.inv_std[0]);
          wq = pk((v.y - p.mean[1]) * p.inv_std[1], (v.w - p.mean[1]) * p.inv_std[1]);
          if (p.Aout) {
            *reinterpret_cast<uint32_t*>(p.Aout + (size_t)(m0 + rowI) * p.K + r * pp + cc) = wi;
            *reinterpret_cast<uint32_t*>(p.Aout + (size_t)(m0 + rowQ) * p.K + r * pp + cc) = wq;
          }
        }
        sts32(sa + a_off(rowI, kcol), wi);
        sts32(sa + a_off(rowQ, kcol), wq);
      }
    } else {
      // model layout [B, 1, 32, 64]: a half-warp reads one image row (64 px), a warp two consecutive selected rows
      const int nphh = 32 / pp, nir = fpt * nphh * rpk;
      for (int wl = cw; 2 * wl < nir; wl += F_NCONV) {
        const int ir = 2 * wl + (lane >> 4);
        const int rr = ir % rpk, t2 = ir / rpk, phh = t2 % nphh, f = t2 / nphh;
        const int r = kb * rpk + rr, x = (lane & 15) * 4, pw = x / pp, cc = x - pw * pp;
        const int row = f * p.Ttok + phh * wp + pw, kcol = rr * pp + cc;
        if (ir < nir) {
          uint32_t w0 = 0u, w1 = 0u;
          if (f0 + f < p.B) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(f0 + f) * 32 + phh * pp + r) * 64 + x));
            w0 = pk(v.x, v.y);
            w1 = pk(v.z, v.w);
            if (p.Aout) *reinterpret_cast<uint2*>(p.Aout + (size_t)(m0 + row) * p.K + r * pp + cc) = make_uint2(w0, w1);
          }
          sts64(sa + a_off(row, kcol), w0, w1);
        }
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(F_THREADS, 1)
frontend_tc_kernel(const __grid_constant__ CUtensorMap mapW, const FrontParams p, const Epi epi,
                   const float* __restrict__ src, int nstage) {
  using C = FrontCfg<BN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sW = smem;                                            // [kb_total][BN x 64] bf16, SWIZZLE_128B
  unsigned char* sA = sW + (size_t)p.kb_total * C::W_BLOCK;           // ring of nstage A k-blocks
  float* stage_base = reinterpret_cast<float*>(sA + (size_t)nstage * C::A_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(stage_base) + C::STG_BYTES);
  uint64_t* wfull = bars;
  uint64_t* full_bar = bars + 1;            // [nstage <= 8]
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tn = blockIdx.x % p.tiles_n, cta_m = blockIdx.x / p.tiles_n, m_stride = gridDim.x / p.tiles_n;
  const int n0 = tn * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW);
    mbar_init(wfull, 1);
    for (int s = 0; s < nstage; ++s) {
      mbar_init(full_bar + s, F_NCONV);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + s, 1);
      mbar_init(tempty_bar + s, F_NEPI);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  // columns of the A tiles beyond K (K % 64 != 0) are never written by the converters: they must read as zero
  for (int i = threadIdx.x; i < nstage * C::A_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, (uint32_t)(p.kb_total * C::W_BLOCK));
      for (int kb = 0; kb < p.kb_total; ++kb) tma_load_2d(&mapW, wfull, sW + (size_t)kb * C::W_BLOCK, kb * FBK, n0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(FBM, BN, 0);
      mbar_wait(wfull, 0);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tm = cta_m; tm < p.tiles_m; tm += m_stride) {
        mbar_wait(tempty_bar + as, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(sA + (size_t)stage * C::A_BYTES);
          const uint32_t sb = smem_u32(sW + (size_t)kb * C::W_BLOCK);
          const int ksteps = min(FBK, p.K - kb * FBK + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_d, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar + stage);
          if (++stage == nstage) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + as);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp < F_EPI_WARP0) {
    // ===================== converter warps =====================
    const int cw = warp - F_CONV_WARP0;
    int stage = 0;
    uint32_t phase = 0;
    for (int tm = cta_m; tm < p.tiles_m; tm += m_stride) {
      for (int kb = 0; kb < p.kb_total; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1);
        convert_block(p, src, smem_u32(sA + (size_t)stage * C::A_BYTES), tm * FBM, kb, cw, lane);
        fence_proxy_async();           // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar + stage);
        if (++stage == nstage) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - F_EPI_WARP0;
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                          // which half of the tile's columns this warp drains
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t stg = smem_u32(stage_base + ew * (32 * F_STG_LD));
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    for (int tm = cta_m; tm < p.tiles_m; tm += m_stride) {
      const int m_base = tm * FBM + quad * 32;
      mbar_wait(tfull_bar + as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
      const int c_begin = half * (BN / 2), c_end = min((half + 1) * (BN / 2), p.d - n0);
      uint32_t r[32];
      if (c_begin < c_end) tmem_ld32(taddr + c_begin, r);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg + (uint32_t)(lane * F_STG_LD + (((j >> 2) ^ (lane & 7)) << 2)) * 4, r[j], r[j + 1], r[j + 2], r[j + 3]);
        if (c + 32 < c_end) tmem_ld32(taddr + c + 32, r);
        __syncwarp();
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + sub_r;
          v[it] = lds128(stg + (uint32_t)(rr * F_STG_LD + ((((lane & 7)) ^ (rr & 7)) << 2)) * 4);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it)
          epi_apply4<bf16>(epi, m_base + it * 4 + sub_r, n0 + c + sub_c, v[it], p.M, p.d, p.vec_ok != 0);
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

int encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    AMC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    AMC_CHECK_ARG(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return 0;
}

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <int BN>
int launch_front(const AmcDesc& D, FrontParams p, const Epi& epi, const float* src, const bf16* W, cudaStream_t st) {
  using C = FrontCfg<BN>;
  EncodeTiledFn enc;
  AMC_TRY(encode_fn(&enc));
  CUtensorMap mapW;
  cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.d};
  cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
  cuuint32_t box[2] = {(cuuint32_t)FBK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(W), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AMC_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (embedding weight) failed (%d)", (int)r);
  p.tiles_n = ceil_div(p.d, BN);
  p.tiles_m = ceil_div(p.M, FBM);
  int nstage = std::min(8, std::max(2, 2 * p.kb_total));
  while (nstage > 2 && C::smem(p.kb_total, nstage) > 227 * 1024) --nstage;
  const size_t smem = C::smem(p.kb_total, nstage);
  AMC_CHECK_ARG(smem <= 227 * 1024, "front end: K=%d d=%d needs %zu bytes of shared memory", p.K, p.d, smem);
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_n = std::max(1, std::min(p.tiles_m, sms / p.tiles_n));
  auto kern = frontend_tc_kernel<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  kern<<<per_n * p.tiles_n, F_THREADS, smem, st>>>(mapW, p, epi, src, nstage);
  AMC_LAUNCH_CHECK();
  (void)D;
  return 0;
}

}  // namespace

// Returns 0 and sets *handled when the geometry is one the fused kernel covers; otherwise the caller runs the
// patchify kernel + generic GEMM.
int frontend_fused(const AmcDesc& D, int Ttok, int K, const float* src, const bf16* W, const Epi& epi, bf16* Aout,
                   bool probe_only, bool* handled, cudaStream_t st) {
  *handled = false;
  const int d = D.d;
  if (D.B <= 0 || d % 8 != 0 || K % 8 != 0 || K > 256 || d > 512) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) != 0) return 0;
  FrontParams p = {};
  p.kind = D.kind; p.raw = D.input_layout == AMC_INPUT_RAW;
  p.B = D.B; p.Ttok = Ttok; p.M = D.B * Ttok; p.K = K; p.d = d;
  p.kb_total = ceil_div(K, FBK);
  if (D.kind == AMC_KIND_RAWIQ) {
    const int S = D.seg;
    // one k-block (2S <= 64); 128-token tiles made of whole 64-sample warp-loads; float4 never straddles a token
    if (D.in_ch != 2 || !pow2(S) || S < 4 || 2 * S > 64 || K != 2 * S) return 0;
    if (!p.raw && (Ttok % 1 != 0 || D.seq_len % 4 != 0)) return 0;
    p.S = S; p.L = D.seq_len;
  } else {
    if (D.in_ch != 1 || D.img_h != 32 || D.img_w != 64 || (D.patch != 4 && D.patch != 8 && D.patch != 16)) return 0;
    if (FBM % Ttok != 0 && Ttok % FBM != 0) return 0;
    p.p = D.patch;
  }
  if (probe_only) {
    *handled = true;
    return 0;
  }
  p.mean[0] = D.norm[0]; p.inv_std[0] = 1.f / D.norm[1];
  p.mean[1] = D.norm[2]; p.inv_std[1] = 1.f / D.norm[3];
  p.vec_ok = epi_vec_ok<bf16>(epi, d) ? 1 : 0;
  p.Aout = Aout;
  if (d <= 128) AMC_TRY(launch_front<128>(D, p, epi, src, W, st));
  else AMC_TRY(launch_front<256>(D, p, epi, src, W, st));
  *handled = true;
  return 0;
}

}  // namespace amc
