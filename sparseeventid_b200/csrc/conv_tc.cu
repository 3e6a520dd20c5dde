// Output-stationary gather-GEMM convolution on 5th-generation tensor cores (sm_100a only):
//     out[o, :] = bias + sum_k  in[nbr[k][o], :] . B_k            (bf16 operands, fp32 accumulate)
// One persistent CTA per SM walks 128-row output tiles.  Per (kernel offset k, 64-channel chunk):
//   * 4 producer warps gather the 128 neighbour rows (16-byte cp.async, zero-fill for missing
//     neighbours) into a 128B-swizzled K-major A tile in shared memory,
//   * 1 thread streams the matching pre-swizzled weight tile B_k with a 1-D bulk async copy (TMA
//     engine, mbarrier complete_tx),
//   * 1 thread issues tcgen05.mma (M=128, N=n_out, K=16) accumulating in TMEM,
//   * 4 epilogue warps drain the finished accumulator (tcgen05.ld), add bias, convert and store the
//     tile while the next tile's MMAs run into the second TMEM buffer.
// Every output row is written exactly once: no atomics, deterministic.
// Replaces SCN's dConvolution_KMxKN_forwardA/B; reference call sites
// src/networks/sparse_building_blocks.py:29-34,110-117.
#include <cstdlib>

#include "common.cuh"

namespace tc {

constexpr int BM = 128;                 // output rows per tile == TMEM lanes
constexpr int KC = 64;                  // channels per pipeline stage (one 128-byte swizzle row)
constexpr int A_BYTES = BM * 128;       // 16 KB
constexpr int EPI_WARPS = 4;            // warps 0..3  (TMEM lane quarter = warp index)
constexpr int PROD_WARPS = 4;           // warps 4..7
constexpr int WARP_MMA = 8;
constexpr int WARP_BLOAD = 9;
constexpr int THREADS = 320;
constexpr int MAX_STAGES = 8;
constexpr int LAG = 3;                  // cp.async groups kept in flight per producer thread
constexpr uint32_t SPIN_LIMIT = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// rows are 128 bytes, groups of 8 rows are 1024 bytes apart (SBO); LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;       // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;       // LayoutType::SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, A and B K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
  const __nv_bfloat16* in;      // [n_in_rows, n_in]
  const int32_t* nbr;           // [K][n_pad]
  const unsigned char* bimg;    // [K][nch][n_out][128 B] pre-swizzled weight tiles
  const float* bias;            // [n_out] or null
  __nv_bfloat16* out;           // [n_rows, n_out]
  int64_t n_rows, n_pad;
  int K, n_in, n_out, nch, last_kc, stages, num_tiles;
};

__global__ void __launch_bounds__(THREADS, 1) k_conv_tc(const Params p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gbase = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t b_bytes = (uint32_t)p.n_out * 128u;
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  const uint32_t bar0 = base + (uint32_t)S * stage_bytes;  // 8-byte aligned (stage_bytes % 1024 == 0)
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(MAX_STAGES + s); };
  auto accf_bar = [&](int b) { return bar0 + 8u * (uint32_t)(2 * MAX_STAGES + b); };
  auto acce_bar = [&](int b) { return bar0 + 8u * (uint32_t)(2 * MAX_STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + (size_t)S * stage_bytes + 8 * (2 * MAX_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(full_bar(s), PROD_WARPS + 1);   // 4 producer warps + the B loader's expect_tx arrival
        mbar_init(empty_bar(s), 1);               // one tcgen05.commit
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf_bar(b), 1);
        mbar_init(acce_bar(b), EPI_WARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int stages_per_tile = p.K * p.nch;

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // ================================ A producers =========================================
    const int r = (warp - EPI_WARPS) * 32 + lane;             // row of the tile owned by this thread
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    uint32_t it = 0;                                           // stages issued by this thread
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t row = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * BM + r;
      int j_next = p.nbr[row];
      for (int k = 0; k < p.K; ++k) {
        const int j = j_next;
        if (k + 1 < p.K) j_next = p.nbr[(int64_t)(k + 1) * p.n_pad + row];
        const __nv_bfloat16* src_row = p.in + (int64_t)(j >= 0 ? j : 0) * p.n_in;
        const int nbytes = j >= 0 ? 16 : 0;
        for (int ch = 0; ch < p.nch; ++ch, ++it) {
          const int s = (int)(it % (uint32_t)S);
          mbar_wait(empty_bar(s), ((it / (uint32_t)S) & 1u) ^ 1u);
          const uint32_t dst = base + (uint32_t)s * stage_bytes + row_off;
          const int nchunk = (ch == p.nch - 1 ? p.last_kc : KC) >> 3;
          const __nv_bfloat16* src = src_row + ch * KC;
#pragma unroll 8
          for (int c = 0; c < nchunk; ++c) cp_async16(dst + (((uint32_t)c ^ sw) << 4), src + c * 8, nbytes);
          cp_async_commit();
          if (it >= (uint32_t)LAG) {
            cp_async_wait<LAG>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar((int)((it - LAG) % (uint32_t)S)));
          }
        }
      }
    }
    // drain: everything issued has landed after wait_group 0
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const uint32_t first = it >= (uint32_t)LAG ? it - LAG : 0u;
      for (uint32_t q = first; q < it; ++q) mbar_arrive(full_bar((int)(q % (uint32_t)S)));
    }
  } else if (warp == WARP_BLOAD) {
    // ================================ B loader (1 thread) ===================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int q = 0; q < stages_per_tile; ++q, ++it) {
          const int s = (int)(it % (uint32_t)S);
          mbar_wait(empty_bar(s), ((it / (uint32_t)S) & 1u) ^ 1u);
          mbar_expect_tx(full_bar(s), b_bytes);
          bulk_g2s(base + (uint32_t)s * stage_bytes + A_BYTES, p.bimg + (size_t)q * b_bytes, b_bytes, full_bar(s));
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer (1 thread) =================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.n_out);
      uint32_t it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait(acce_bar(buf), (((uint32_t)t >> 1) & 1u) ^ 1u);     // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u;
        for (int q = 0; q < stages_per_tile; ++q, ++it) {
          const int s = (int)(it % (uint32_t)S);
          mbar_wait(full_bar(s), (it / (uint32_t)S) & 1u);
          tc_fence_after();
          const uint32_t a_addr = base + (uint32_t)s * stage_bytes;
          const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + A_BYTES);
          const int ch = q % p.nch;
          const int nk = (ch == p.nch - 1 ? p.last_kc : KC) >> 4;
          for (int kk = 0; kk < nk; ++kk)
            umma(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (q > 0 || kk > 0) ? 1u : 0u);
          umma_commit(empty_bar(s));                                   // frees the stage when the MMAs retire
        }
        umma_commit(accf_bar(buf));                                    // accumulator complete
      }
    }
  } else {
    // ================================ epilogue (warps 0..3) ==================================
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(accf_bar(buf), ((uint32_t)t >> 1) & 1u);
      tc_fence_after();
      const int64_t row = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * BM + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * 256u;
      for (int c0 = 0; c0 < p.n_out; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (row < p.n_rows) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.n_out + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g * 8 + e]) + (p.bias ? __ldg(p.bias + c0 + g * 8 + e) : 0.f);
            uint4 u;
            u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
            u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
            dst[g] = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Weight image for the kernel above: tile (k, ch) = n_out rows of 128 bytes, 16-byte chunk c of row n
// stored at chunk position c ^ (n & 7)  (the SWIZZLE_128B pattern); unused half rows stay zero.
__global__ void k_prep_weights_tc(const float* __restrict__ W, int K, int Cin, int Cout, int transpose, int mirror,
                                  int nch, __nv_bfloat16* __restrict__ img) {
  const int n_in = transpose ? Cout : Cin, n_out = transpose ? Cin : Cout;
  int64_t total = (int64_t)K * n_in * n_out;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i / ((int64_t)n_in * n_out));
  int rem = (int)(i - (int64_t)k * n_in * n_out);
  int n = rem / n_in, c = rem % n_in;
  int src_k = (transpose && mirror) ? K - 1 - k : k;
  int ci = transpose ? n : c, co = transpose ? c : n;
  float v = W[((int64_t)src_k * Cin + ci) * Cout + co];
  int ch = c / KC, cc = c % KC;
  size_t off = (((size_t)k * nch + ch) * n_out + n) * 64 + (size_t)(((cc >> 3) ^ (n & 7)) << 3) + (cc & 7);
  img[off] = __float2bfloat16_rn(v);
}

}  // namespace tc

bool scn_tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SCN_B200_DISABLE_TC");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

bool scn_tc_shape_ok(int K, int n_in, int n_out) {
  return K >= 1 && (n_in % 32) == 0 && (n_out % 32) == 0 && n_in >= 32 && n_in <= 256 && n_out >= 32 && n_out <= 256;
}

size_t scn_tc_image_bytes(int K, int n_in, int n_out) {
  int nch = (n_in + tc::KC - 1) / tc::KC;
  return (size_t)K * nch * n_out * 128;
}

int scn_tc_prep(const float* W, int K, int Cin, int Cout, int transpose, int mirror, void* out, cudaStream_t s) {
  const int n_in = transpose ? Cout : Cin, n_out = transpose ? Cin : Cout;
  const int nch = (n_in + tc::KC - 1) / tc::KC;
  if (n_in % tc::KC) SCN_CUDA(cudaMemsetAsync(out, 0, scn_tc_image_bytes(K, n_in, n_out), s));
  int64_t total = (int64_t)K * Cin * Cout;
  tc::k_prep_weights_tc<<<grid_for(total, 256), 256, 0, s>>>(W, K, Cin, Cout, transpose, mirror, nch,
                                                             (__nv_bfloat16*)out);
  SCN_LAUNCH_CHECK();
  return SCN_OK;
}

int scn_tc_forward(const __nv_bfloat16* in, const int32_t* nbr, int K, int64_t n_rows, int64_t n_pad, int n_in,
                   int n_out, const void* bimg, const float* bias, __nv_bfloat16* out, cudaStream_t s) {
  tc::Params p;
  p.in = in; p.nbr = nbr; p.bimg = (const unsigned char*)bimg; p.bias = bias; p.out = out;
  p.n_rows = n_rows; p.n_pad = n_pad; p.K = K; p.n_in = n_in; p.n_out = n_out;
  p.nch = (n_in + tc::KC - 1) / tc::KC;
  p.last_kc = n_in - (p.nch - 1) * tc::KC;
  const uint32_t stage_bytes = tc::A_BYTES + (uint32_t)n_out * 128u;
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
  if (stages < tc::LAG + 1) return SCN_ERR_UNSUPPORTED;
  p.stages = stages;
  p.num_tiles = (int)((n_rows + tc::BM - 1) / tc::BM);
  size_t smem = 1024 + (size_t)stages * stage_bytes + 8 * (2 * tc::MAX_STAGES + 4) + 16;
  SCN_CUDA(cudaFuncSetAttribute(tc::k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  tc::k_conv_tc<<<grid, tc::THREADS, smem, s>>>(p);
  SCN_LAUNCH_CHECK();
  return SCN_OK;
}
