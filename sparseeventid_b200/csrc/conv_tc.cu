// Output-stationary gather-GEMM convolution on 5th-generation tensor cores (sm_100a only):
//     out[o, :] = bias + sum_k  in[nbr[k][o], :] . B_k            (bf16 operands, fp32 accumulate)
//
// One persistent CTA per SM walks GROUPS of T 128-row output tiles (T * n_out <= 256 TMEM columns,
// two groups double-buffered in the 512 columns).  The contraction (K offsets x n_in channels) is cut
// into stages of 64 channels (one 128-byte shared-memory row).  For every stage q = (offset, chunk):
//   * the weight tile B(q) (n_out x 128 B, pre-swizzled image) is streamed ONCE per group by a 1-D bulk
//     async copy (TMA engine, mbarrier complete_tx) and reused by the T tiles of the group; the group's
//     T*128 neighbour indices arrive the same way into an index ring,
//   * for each tile, 8 producer warps gather the 128 neighbour rows into a 128B-swizzled K-major A tile:
//     8 lanes move one 128-byte row (coalesced 16-byte LDG), the loads of D stages are in flight per thread
//     before the first is stored (STS.128), missing neighbours are written as zeros without a global read,
//   * 1 thread issues tcgen05.mma (M=128, N=n_out, K=16) accumulating in TMEM,
//   * 4 epilogue warps drain finished accumulators (tcgen05.ld), add bias, convert and store while the
//     next group's MMAs run into the other TMEM half.
// Every output row is written exactly once: no atomics, deterministic.
//
// Measured alternatives for the A gather (profiles/): 16-byte cp.async tops out near
// 16 B/clk/SM (~950 cycles per 16 KB stage); TMA tile::gather4 (kept as an option, SCN_B200_TC_GATHER=tma)
// costs ~77 cycles per 512-byte instruction.  LDG.128 + STS.128 is the fastest of the three.
//
// Replaces SCN's dConvolution_KMxKN_forwardA/B; reference call sites
// src/networks/sparse_building_blocks.py:29-34,110-117.
#include <cuda.h>      // CUtensorMap + enums only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace tc {

constexpr int BM = 128;                 // output rows per tile == TMEM lanes
constexpr int KC = 64;                  // channels per pipeline stage (one 128-byte swizzle row)
constexpr int A_BYTES = BM * 128;       // 16 KB
constexpr int EPI_WARPS = 4;            // warps 0..3  (TMEM lane quarter = warp index)
constexpr int PROD_WARPS = 8;           // warps 4..11
constexpr int WARP_MMA = 12;
constexpr int WARP_BLOAD = 13;          // weight tiles
constexpr int WARP_ILOAD = 14;          // neighbour-index blocks (separate thread: must never wait on the B ring)
constexpr int THREADS = 480;
constexpr int MAX_A = 12, MAX_B = 6, MAX_I = 4;   // ring depths: A tiles, B tiles, neighbour-index blocks
constexpr int D = 4;                    // stages whose global loads a producer thread keeps in flight
constexpr int TMA_LANES = 4;            // TMA mode: lanes 0..3 of each producer warp issue one gather4 per stage
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
// long waits (epilogue warps idle through a whole group's main loop): back off so the spin does not
// compete with the producer warps for issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(256);
    if (++spins > SPIN_LIMIT) __trap();
  }
}
// Warp-uniform leader election.  The issuing warps keep their whole control flow warp-uniform and only predicate
// the tcgen05 / bulk-copy instruction itself on the elected lane: operands then live in uniform registers.  (With an
// `if (lane == 0)` around the loop the compiler cannot prove uniformity and wraps every UTCHMMA / UTCBAR / UBLKCP in
// an R2UR + ELECT + BRA.U.ANY waterfall: ~80 cycles each, which made the single MMA thread the bottleneck.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint4 ldg_nc128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// TMA tile::gather4: four rows (arbitrary row coordinates r0..r3, one box of 64 channels starting at column c)
// of a 2-D tensor land as four consecutive 128-byte rows at dst, swizzled by the tensor map (SWIZZLE_128B);
// rows/columns outside the tensor are zero-filled without touching memory.  512 bytes complete_tx on `bar`.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int c, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
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
  const unsigned char* bimg;    // [K*nch][n_out][128 B] pre-swizzled weight tiles, one per stage
  const float* bias;            // [n_out] or null
  __nv_bfloat16* out;           // [n_rows, n_out]
  int64_t n_rows, n_pad;
  int K, n_in, n_out;
  int last_kc;                  // channels in the last 64-channel chunk of an offset (64 or 32)
  int T;                        // tiles per group
  int SA, SB;                   // A / B ring depth
  int num_tiles, num_groups;
  int use_tma;                  // A tiles by TMA gather4 (1) or by LDG+STS (0)
  int n_in_rows;                // rows of `in` (gather4: any row index >= n_in_rows is zero-filled)
};

// NCH: 64-channel chunks per offset = ceil(n_in / 64).  PAIR (n_in == 32, NCH == 1): one stage holds TWO offsets,
// 32 channels each (chunks 0-3 from offset 2q, chunks 4-7 from offset 2q+1), halving the stage count.
template <int NCH, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) k_conv_tc(const Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gbase = smem_raw + (base - raw);
  const int SA = p.SA, SB = p.SB;
  const uint32_t b_bytes = (uint32_t)p.n_out * 128u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + (uint32_t)SA * A_BYTES;
  const uint32_t i_bytes = (uint32_t)(PAIR ? 2 : 1) * (uint32_t)p.T * 512u;   // index block: T tiles x 128 rows (x2 offsets)
  const uint32_t i_base = b_base + (uint32_t)SB * b_bytes;
  const uint32_t bar0 = i_base + (uint32_t)MAX_I * i_bytes;  // 8-byte aligned
  auto afull = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto aempty = [&](int s) { return bar0 + 8u * (uint32_t)(MAX_A + s); };
  auto bfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 * MAX_A + s); };
  auto bempty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * MAX_A + MAX_B + s); };
  auto ifull = [&](int s) { return bar0 + 8u * (uint32_t)(2 * MAX_A + 2 * MAX_B + s); };
  auto iempty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * MAX_A + 2 * MAX_B + MAX_I + s); };
  auto accf = [&](int b) { return bar0 + 8u * (uint32_t)(2 * MAX_A + 2 * MAX_B + 2 * MAX_I + b); };
  auto acce = [&](int b) { return bar0 + 8u * (uint32_t)(2 * MAX_A + 2 * MAX_B + 2 * MAX_I + 2 + b); };
  constexpr int NBAR = 2 * MAX_A + 2 * MAX_B + 2 * MAX_I + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + (size_t)SA * A_BYTES + (size_t)SB * b_bytes +
                                                    (size_t)MAX_I * i_bytes + 8 * NBAR);
  const int* sidx_all = reinterpret_cast<const int*>(gbase + (size_t)SA * A_BYTES + (size_t)SB * b_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < SA; ++s) {
        // LDG/STS mode: one arrival per producer warp once its rows are stored and fenced;
        // TMA mode: one arrive.expect_tx(512) per issuing lane, completed by the gather4 bytes
        mbar_init(afull(s), p.use_tma ? PROD_WARPS * TMA_LANES : PROD_WARPS);
        mbar_init(aempty(s), 1);                  // one tcgen05.commit
      }
      for (int s = 0; s < SB; ++s) {
        mbar_init(bfull(s), 1);                   // the loader's expect_tx arrival (+ complete_tx bytes)
        mbar_init(bempty(s), 1);
      }
      for (int s = 0; s < MAX_I; ++s) {
        mbar_init(ifull(s), 1);
        mbar_init(iempty(s), PROD_WARPS);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf(b), 1);
        mbar_init(acce(b), EPI_WARPS);
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

  const int my_groups = (p.num_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int T = p.T;
  const int NSTEP = PAIR ? (p.K + 1) / 2 : p.K;            // index blocks per tile
  const int Q = NSTEP * NCH;                                // stages per tile

  if (!PAIR && warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS && p.use_tma) {
    // ================================ A producers, TMA gather4 (option) =====================
    // lane (pw, l < 4) owns tile rows [16pw + 4l, +4) and moves them with ONE gather4 per stage.  Missing
    // neighbours (-1) become an out-of-range row index, which the TMA engine zero-fills without reading.
    const int pw = warp - EPI_WARPS;
    if (lane < TMA_LANES) {
      const int r0 = pw * 16 + lane * 4;
      const uint32_t dst_off = (uint32_t)r0 * 128u;
      const int oob = p.n_in_rows;
      int slot = 0, islot = 0;
      uint32_t round = 0, iround = 0;
      for (int g = 0; g < my_groups; ++g) {
        const int64_t tile0 = ((int64_t)blockIdx.x + (int64_t)g * gridDim.x) * T;
        const int tvalid = (int)((int64_t)p.num_tiles - tile0 < (int64_t)T ? (int64_t)p.num_tiles - tile0 : (int64_t)T);
        for (int step = 0; step < p.K; ++step) {
          mbar_wait(ifull(islot), iround & 1u);
          const int* sidx = sidx_all + (size_t)islot * (i_bytes / 4);
#pragma unroll
          for (int chn = 0; chn < NCH; ++chn) {
            for (int t = 0; t < T; ++t) {
              int4 jj = make_int4(-1, -1, -1, -1);
              if (t < tvalid) jj = *reinterpret_cast<const int4*>(sidx + t * 128 + r0);
              jj.x = jj.x < 0 ? oob : jj.x; jj.y = jj.y < 0 ? oob : jj.y;
              jj.z = jj.z < 0 ? oob : jj.z; jj.w = jj.w < 0 ? oob : jj.w;
              mbar_wait(aempty(slot), (round & 1u) ^ 1u);
              mbar_expect_tx(afull(slot), 512u);
              tma_gather4(a_base + (uint32_t)slot * A_BYTES + dst_off, &tmap, chn * KC, jj.x, jj.y, jj.z, jj.w, afull(slot));
              if (++slot == SA) { slot = 0; ++round; }
            }
          }
          // the index block is released by lane 0 of every producer warp (count = PROD_WARPS)
          __syncwarp(0xFu);
          if (lane == 0) mbar_arrive(iempty(islot));
          if (++islot == MAX_I) { islot = 0; ++iround; }
        }
      }
    }
  } else if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // ================================ A producers, LDG.128 -> STS.128 ========================
    // Warp pw owns tile rows [16pw, 16pw+16).  In pass i (0..3) the 8 lanes with the same (lane >> 3) move one
    // whole 128-byte row: lane handles 16-byte chunk (lane & 7) of row 16pw + 4i + (lane >> 3), so a warp-wide
    // load touches 4 contiguous 128-byte lines.  Work is done in batches of D stages: first all global loads of
    // the batch are issued (the neighbour indices come from the shared-memory ring; no waiting on free A slots),
    // then each stage is stored to its slot as soon as the slot is free, fenced for the async proxy (tensor
    // core reads) and signalled with ONE mbarrier arrival per warp.
    const int pw = warp - EPI_WARPS;
    const int chunk = lane & 7;
    const int sub = lane >> 3;
    uint32_t dst_off[4];
    int rowi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rt = pw * 16 + i * 4 + sub;                    // row within the tile
      rowi[i] = rt;
      dst_off[i] = (uint32_t)rt * 128u + (((uint32_t)chunk ^ (uint32_t)(rt & 7)) << 4);
    }
    const bool hi_half = PAIR && chunk >= 4;                   // this lane's 16 bytes belong to offset 2q+1
    const uint32_t col_bytes = PAIR ? (uint32_t)(chunk & 3) * 16u : (uint32_t)chunk * 16u;
    const uint32_t row_bytes = (uint32_t)p.n_in * 2u;
    const unsigned char* in_bytes = reinterpret_cast<const unsigned char*>(p.in);
    const int last_chunks = p.last_kc >> 3;

    // issue cursor over (group, offset, chunk, tile)
    int cg = 0, ck = 0, cc = 0, ct = 0;
    int tvalid;
    {
      const int64_t tile0 = (int64_t)blockIdx.x * T;
      tvalid = (int)((int64_t)p.num_tiles - tile0 < (int64_t)T ? (int64_t)p.num_tiles - tile0 : (int64_t)T);
    }
    int islot = 0, slot = 0;
    uint32_t iround = 0, round = 0;
    int64_t remaining = (int64_t)my_groups * Q * T;
    while (remaining > 0) {
      const int nb = remaining < D ? (int)remaining : D;
      uint4 v[D][4];
      // ---- phase 1: issue the global loads of up to D stages -------------------------------------------
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (d < nb) {
          if (cc == 0 && ct == 0) mbar_wait(ifull(islot), iround & 1u);      // first stage of an offset
          const int* sidx = sidx_all + (size_t)islot * (i_bytes / 4) + ct * 128 + (hi_half ? T * 128 : 0);
          const bool lane_on = PAIR ? true : chunk < (cc == NCH - 1 ? last_chunks : 8);
          const uint32_t coff = col_bytes + (uint32_t)cc * 128u;
          const bool tile_ok = ct < tvalid && !(hi_half && 2 * ck + 1 >= p.K);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = tile_ok ? sidx[rowi[i]] : -1;
            v[d][i] = make_uint4(0u, 0u, 0u, 0u);
            if (j >= 0 && lane_on) v[d][i] = ldg_nc128(in_bytes + ((size_t)(uint32_t)j * row_bytes + coff));
          }
          // advance the cursor; when an offset is finished release its index block
          if (++ct == T) {
            ct = 0;
            if (++cc == NCH) {
              cc = 0;
              __syncwarp();
              if (lane == 0) mbar_arrive(iempty(islot));
              if (++islot == MAX_I) { islot = 0; ++iround; }
              if (++ck == NSTEP) {
                ck = 0;
                ++cg;
                const int64_t tile0 = ((int64_t)blockIdx.x + (int64_t)cg * gridDim.x) * T;
                tvalid = (int)((int64_t)p.num_tiles - tile0 < (int64_t)T ? (int64_t)p.num_tiles - tile0 : (int64_t)T);
              }
            }
          }
        }
      }
      // ---- phase 2: store each stage into its A slot as soon as the slot is free ... -------------------------
      const int slot0 = slot;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (d < nb) {
          mbar_wait(aempty(slot), (round & 1u) ^ 1u);
          const uint32_t abase = a_base + (uint32_t)slot * A_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) sts128(abase + dst_off[i], v[d][i]);
          if (++slot == SA) { slot = 0; ++round; }
        }
      }
      // ---- ... then ONE proxy fence + warp sync for the whole batch, and one arrival per stage ------------
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        int sl = slot0;
        for (int d = 0; d < nb; ++d) {
          mbar_arrive(afull(sl));
          if (++sl == SA) sl = 0;
        }
      }
      remaining -= nb;
    }
  } else if (warp == WARP_ILOAD) {
    // ================================ neighbour-index loader (1 elected lane) ===============
    int islot = 0;
    uint32_t iround = 0;
    for (int g = 0; g < my_groups; ++g) {
      const int64_t tile0 = ((int64_t)blockIdx.x + (int64_t)g * gridDim.x) * T;
      const int tvalid = (int)((int64_t)p.num_tiles - tile0 < (int64_t)T ? (int64_t)p.num_tiles - tile0 : (int64_t)T);
      const uint32_t blk = (uint32_t)tvalid * 512u;
      for (int step = 0; step < NSTEP; ++step) {
        // indices of the group's tvalid*128 consecutive rows for this offset (PAIR: offsets 2*step and 2*step+1)
        mbar_wait(iempty(islot), (iround & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t idst = i_base + (uint32_t)islot * i_bytes;
          if (PAIR) {
            const bool two = 2 * step + 1 < p.K;
            mbar_expect_tx(ifull(islot), two ? 2u * blk : blk);
            bulk_g2s(idst, p.nbr + (int64_t)(2 * step) * p.n_pad + tile0 * BM, blk, ifull(islot));
            if (two)
              bulk_g2s(idst + (uint32_t)T * 512u, p.nbr + (int64_t)(2 * step + 1) * p.n_pad + tile0 * BM, blk, ifull(islot));
          } else {
            mbar_expect_tx(ifull(islot), blk);
            bulk_g2s(idst, p.nbr + (int64_t)step * p.n_pad + tile0 * BM, blk, ifull(islot));
          }
        }
        __syncwarp();
        if (++islot == MAX_I) { islot = 0; ++iround; }
      }
    }
  } else if (warp == WARP_BLOAD) {
    // ================================ weight-tile loader (1 elected lane) ===================
    int bslot = 0;
    uint32_t bround = 0;
    for (int g = 0; g < my_groups; ++g) {
      for (int q = 0; q < Q; ++q) {
        mbar_wait(bempty(bslot), (bround & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(bfull(bslot), b_bytes);
          bulk_g2s(b_base + (uint32_t)bslot * b_bytes, p.bimg + (size_t)q * b_bytes, b_bytes, bfull(bslot));
        }
        __syncwarp();
        if (++bslot == SB) { bslot = 0; ++bround; }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer (warp-uniform loop, 1 elected lane issues) ==
    const uint32_t idesc = make_idesc(p.n_out);
    int aslot = 0, bslot = 0;
    uint32_t around = 0, bround = 0;
    for (int g = 0; g < my_groups; ++g) {
      const int buf = g & 1;
      mbar_wait(acce(buf), (((uint32_t)g >> 1) & 1u) ^ 1u);         // epilogue has drained this TMEM half
      tc_fence_after();
      for (int q = 0; q < Q; ++q) {
        mbar_wait(bfull(bslot), bround & 1u);
        const uint64_t db = make_desc_sw128(b_base + (uint32_t)bslot * b_bytes);
        const int nk = PAIR ? 4 : (((q % NCH) == NCH - 1 ? p.last_kc : KC) >> 4);
        for (int t = 0; t < T; ++t) {
          mbar_wait(afull(aslot), around & 1u);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(a_base + (uint32_t)aslot * A_BYTES);
          const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u + (uint32_t)(t * p.n_out);
          if (elect_one()) {
            for (int kk = 0; kk < nk; ++kk)
              umma(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (q > 0 || kk > 0) ? 1u : 0u);
            umma_commit(aempty(aslot));                                // frees the A slot when these MMAs retire
          }
          __syncwarp();
          if (++aslot == SA) { aslot = 0; ++around; }
        }
        if (elect_one()) umma_commit(bempty(bslot));
        __syncwarp();
        if (++bslot == SB) { bslot = 0; ++bround; }
      }
      if (elect_one()) umma_commit(accf(buf));                         // the group's accumulators are complete
      __syncwarp();
    }
  } else if (warp < EPI_WARPS) {
    // ================================ epilogue (warps 0..3) ==================================
    for (int g = 0; g < my_groups; ++g) {
      const int buf = g & 1;
      mbar_wait_sleep(accf(buf), ((uint32_t)g >> 1) & 1u);
      tc_fence_after();
      for (int t = 0; t < T; ++t) {
        const int64_t tile = ((int64_t)blockIdx.x + (int64_t)g * gridDim.x) * T + t;
        if (tile >= p.num_tiles) break;
        const int64_t row = tile * BM + warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * 256u + (uint32_t)(t * p.n_out);
        for (int c0 = 0; c0 < p.n_out; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
          if (row < p.n_rows) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.n_out + c0);
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                f[e] = __uint_as_float(v[gq * 8 + e]) + (p.bias ? __ldg(p.bias + c0 + gq * 8 + e) : 0.f);
              uint4 u;
              u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
              u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
              dst[gq] = u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Weight image: one tile per stage q = k*nch + c/64: n_out rows of 128 bytes; the 16-byte chunk (c%64)/8 of row n
// is stored at chunk position chunk ^ (n & 7) (the SWIZZLE_128B pattern); unused half rows stay zero.
__global__ void k_prep_weights_tc(const float* __restrict__ W, int K, int Cin, int Cout, int transpose, int mirror,
                                  int nch, int pair, __nv_bfloat16* __restrict__ img) {
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
  // pair (n_in == 32): stage q = k/2, chunks 0-3 <- offset 2q, chunks 4-7 <- offset 2q+1
  const int q = pair ? (k >> 1) : k * nch + c / KC;
  const int chunk = pair ? (k & 1) * 4 + (c >> 3) : (c % KC) >> 3;
  size_t off = ((size_t)q * n_out + n) * 64 + (size_t)((chunk ^ (n & 7)) << 3) + (c & 7);
  img[off] = __float2bfloat16_rn(v);
}

}  // namespace tc

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tc_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}
// 0: LDG.128 + STS.128 gather (default), 1: TMA tile::gather4 (SCN_B200_TC_GATHER=tma).
static int tc_gather_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SCN_B200_TC_GATHER");
    v = (e && std::strcmp(e, "tma") == 0) ? 1 : 0;
  }
  return v;
}

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

// n_in == 32: two offsets share a 64-channel stage (not available with the TMA gather, whose box is one row wide)
static bool tc_pair(int n_in) { return n_in == 32 && tc_gather_mode() == 0; }

size_t scn_tc_image_bytes(int K, int n_in, int n_out) {
  const size_t stages = tc_pair(n_in) ? (size_t)(K + 1) / 2 : (size_t)K * ((n_in + tc::KC - 1) / tc::KC);
  return stages * n_out * 128;
}

int scn_tc_prep(const float* W, int K, int Cin, int Cout, int transpose, int mirror, void* out, cudaStream_t s) {
  const int n_in = transpose ? Cout : Cin, n_out = transpose ? Cin : Cout;
  const int nch = (n_in + tc::KC - 1) / tc::KC;
  if ((n_in % tc::KC) != 0) SCN_CUDA(cudaMemsetAsync(out, 0, scn_tc_image_bytes(K, n_in, n_out), s));
  int64_t total = (int64_t)K * Cin * Cout;
  tc::k_prep_weights_tc<<<grid_for(total, 256), 256, 0, s>>>(W, K, Cin, Cout, transpose, mirror, nch,
                                                             tc_pair(n_in) ? 1 : 0, (__nv_bfloat16*)out);
  SCN_LAUNCH_CHECK();
  return SCN_OK;
}

int scn_tc_forward(const __nv_bfloat16* in, int64_t n_in_rows, const int32_t* nbr, int K, int64_t n_rows,
                   int64_t n_pad, int n_in, int n_out, const void* bimg, const float* bias, __nv_bfloat16* out,
                   cudaStream_t s) {
  tc::Params p;
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  p.use_tma = 0;
  p.n_in_rows = (int)n_in_rows;
  if (tc_gather_mode() == 1 && tc_encoder() != nullptr && n_in_rows > 0 && n_in_rows < 0x7fffffffLL) {
    const cuuint64_t gdim[2] = {(cuuint64_t)n_in, (cuuint64_t)n_in_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)n_in * 2u};
    const cuuint32_t box[2] = {(cuuint32_t)tc::KC, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = tc_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(in), gdim, gstride,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SCN_ERR_UNSUPPORTED;
    p.use_tma = 1;
  }
  p.in = in; p.nbr = nbr; p.bimg = (const unsigned char*)bimg; p.bias = bias; p.out = out;
  p.n_rows = n_rows; p.n_pad = n_pad; p.K = K; p.n_in = n_in; p.n_out = n_out;
  const int nch = (n_in + tc::KC - 1) / tc::KC;
  p.last_kc = n_in - (nch - 1) * tc::KC;
  p.num_tiles = (int)((n_rows + tc::BM - 1) / tc::BM);
  // tiles per group: as many accumulators as fit in half of TMEM (weight reuse), chosen to minimise the
  // number of tiles walked by the busiest CTA
  int tmax = 256 / n_out;
  if (tmax < 1) tmax = 1;
  int T = 1;
  long best = -1;
  for (int cand = tmax; cand >= 1; --cand) {
    long groups = (p.num_tiles + cand - 1) / cand;
    long busiest = ((groups + kNumSMs - 1) / kNumSMs) * cand;
    if (best < 0 || busiest < best) { best = busiest; T = cand; }
  }
  p.T = T;
  p.num_groups = (p.num_tiles + T - 1) / T;
  const uint32_t b_bytes = (uint32_t)n_out * 128u;
  // bulk copies have ~1-1.5 us latency: keep enough weight tiles in flight to cover it, within ~72 KB
  {
    int sb = (int)((48u * 1024u) / b_bytes);
    if (sb > tc::MAX_B) sb = tc::MAX_B;
    if (sb < 2) sb = 2;
    p.SB = sb;
  }
  const bool pair = tc_pair(n_in);
  const uint32_t i_bytes = (uint32_t)(pair ? 2 : 1) * (uint32_t)T * 512u;
  constexpr int NBAR = 2 * tc::MAX_A + 2 * tc::MAX_B + 2 * tc::MAX_I + 4;
  const uint32_t fixed = 1024u + (uint32_t)p.SB * b_bytes + (uint32_t)tc::MAX_I * i_bytes + 8u * NBAR + 16u;
  const uint32_t budget = 220u * 1024u;
  int SA = (int)((budget - fixed) / tc::A_BYTES);
  if (SA > tc::MAX_A) SA = tc::MAX_A;
  if (SA < tc::D + 1) return SCN_ERR_UNSUPPORTED;
  p.SA = SA;
  size_t smem = (size_t)fixed + (size_t)SA * tc::A_BYTES;
  int grid = p.num_groups < kNumSMs ? p.num_groups : kNumSMs;
  auto launch = [&](auto kern) -> int {
    SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, tc::THREADS, smem, s>>>(p, tmap);
    SCN_LAUNCH_CHECK();
    return SCN_OK;
  };
  if (pair) return launch(tc::k_conv_tc<1, true>);
  switch (nch) {
    case 1: return launch(tc::k_conv_tc<1, false>);
    case 2: return launch(tc::k_conv_tc<2, false>);
    case 3: return launch(tc::k_conv_tc<3, false>);
    case 4: return launch(tc::k_conv_tc<4, false>);
    default: return SCN_ERR_UNSUPPORTED;
  }
}
