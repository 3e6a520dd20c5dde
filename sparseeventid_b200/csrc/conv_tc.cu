// Output-stationary gather-GEMM convolution on 5th-generation tensor cores (sm_100a only):
//     out[o, :] = bias + sum_k  in[nbr[k][o], :] . B_k            (bf16 operands, fp32 accumulate)
//
// One persistent CTA per SM owns a contiguous range of 128-row output tiles (an even split of the tiles over the
// grid) and walks it in GROUPS of T tiles whose accumulators live in TMEM together (T * n_out <= 256 columns with two
// groups double-buffered, or <= 512 columns single-buffered when a CTA has one group anyway).  The contraction
// (K offsets x n_in channels) is cut into stages of 64 channels (one 128-byte shared-memory row).  Per stage:
//   * the weight tile B(q) (n_out x 128 B, pre-swizzled image) is streamed ONCE per group by a 1-D bulk async copy
//     (TMA engine, mbarrier complete_tx) and reused by the T tiles of the group,
//   * for each tile ONE producer warp gathers the A tile: it reads the tile's 128 neighbour indices (prefetched one
//     stage ahead), ballots them, compacts the LIVE rows into a warp-private list of (source offset, swizzled
//     destination) pairs and copies only those rows with 16-byte LDGSTS (8 lanes per 128-byte row).  The ballots are
//     published as the stage's disable-output-lane mask.  The copies signal their landing themselves
//     (cp.async.mbarrier.arrive.noinc), the warp never waits for them and alternates between its two A slots,
//   * up to 4 issuing warps (tile t -> warp t mod NM) issue tcgen05.mma (M=128, N=n_out, K=16) with that mask:
//     accumulator rows without a neighbour at this offset are not updated, so their stale A rows never reach a
//     result.  Only the first stage of a tile, which initialises the accumulator, is written in full,
//   * 4 epilogue warps drain finished accumulators (tcgen05.ld), add bias, convert and store.
// Every output row is written exactly once: no atomics, deterministic.
//
// Synchronisation notes.  A slot's mbarriers must see consecutive phases from each waiter (a parity wait cannot tell
// phase r from r-2), hence one producer warp per pair of slots; the issuing warps run up to SA stages apart, so the
// "which stage is in this slot" handshake is a monotonically increasing sequence flag (release store / acquire poll)
// and only then the landing barrier, whose parity the flag carries.
//
// What bounds it (timeline marks, -DSCN_TC_TIMELINE + tools/tc_timeline.py): neither L2 latency
// nor the tensor pipe but the dependent-issue rate of the single warps that build a stage (~5 cycles per
// instruction), so every role's per-stage instruction count was cut (6 instructions per gathered row segment,
// ~70 per issued stage) and the roles were multiplied (5-6 producer warps, 4 issuing warps).
// History (profiles/): dense zero-filled A tiles via cp.async, TMA tile::gather4 and lock-step LDG/STS batches
// were measured first; they moved all 128 rows per stage although ~30% are live.
//
// Replaces SCN's dConvolution_KMxKN_forwardA/B; reference call sites
// src/networks/sparse_building_blocks.py:29-34,110-117.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tc {

constexpr int BM = 128;                 // output rows per tile == TMEM lanes
constexpr int KC = 64;                  // channels per pipeline stage (one 128-byte swizzle row)
constexpr int A_BYTES = BM * 128;       // 16 KB
constexpr int EPI_WARPS = 4;            // warps 0..3  (TMEM lane quarter = warp index)
constexpr int PROD_WARPS = 6;           // warps 4..9; the first SA/2 of them are active, two A slots each
constexpr int WARP_MMA = EPI_WARPS + PROD_WARPS;   // first of MMA_WARPS issuing warps (tile t -> warp t mod NM)
constexpr int MMA_WARPS = 4;
constexpr int WARP_BLOAD = WARP_MMA + MMA_WARPS;   // weight tiles
constexpr int THREADS = 32 * (WARP_BLOAD + 1);
constexpr int MAX_A = 2 * PROD_WARPS, MAX_B = 6;   // ring depths: A tiles, B tiles
constexpr int MASK_BYTES = 32;          // per A slot: 2 x 128-bit disable-output-lane masks (second: PAIR upper half)
constexpr int LIST_BYTES = 128 * 8;     // per producer warp: live items of its stage, (source row, smem address); x2 in PAIR mode
using namespace tcptx;

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
  int SA, SB;                   // A / B ring depth (SA == number of active producer warps)
  int num_tiles, num_groups;
  int NM;                       // active MMA-issuing warps = min(T, MMA_WARPS)
  int nbuf;                     // 2: groups alternate between the TMEM halves (T*n_out <= 256); 1: one group uses all 512 columns
  unsigned long long* dbg;      // optional timeline buffer (scn_tc_debug_timeline): CTA 0 records clock64() marks
  int exp;                      // SCN_B200_TC_EXP timing experiments (WRONG results): 2 no MMAs
};

// NCH: 64-channel chunks per offset = ceil(n_in / 64).  PAIR (n_in == 32, NCH == 1): one stage holds TWO offsets,
// 32 channels each (chunks 0-3 from offset 2q, chunks 4-7 from offset 2q+1), halving the stage count.
template <int NCH, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) k_conv_tc(const Params p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gbase = smem_raw + (base - raw);
  const int SA = p.SA, SB = p.SB;
  const uint32_t b_bytes = (uint32_t)p.n_out * 128u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + (uint32_t)SA * A_BYTES;
  const uint32_t bar0 = b_base + (uint32_t)SB * b_bytes;   // 8-byte aligned
  auto afull = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto aempty = [&](int s) { return bar0 + 8u * (uint32_t)(MAX_A + s); };
  auto bfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 * MAX_A + s); };
  auto bempty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * MAX_A + MAX_B + s); };
  auto accf = [&](int b) { return bar0 + 8u * (uint32_t)(2 * MAX_A + 2 * MAX_B + b); };
  auto acce = [&](int b) { return bar0 + 8u * (uint32_t)(2 * MAX_A + 2 * MAX_B + 2 + b); };
  constexpr int NBAR = 2 * MAX_A + 2 * MAX_B + 4;          // 38 -> 304 bytes (a multiple of 16: amask stays 16-byte aligned)
  unsigned char* tail = gbase + (size_t)SA * A_BYTES + (size_t)SB * b_bytes + 8 * NBAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail);
  uint32_t* amask = reinterpret_cast<uint32_t*>(tail + 16);                    // [MAX_A][8], 16-byte aligned
  unsigned char* lists = tail + 16 + MAX_A * MASK_BYTES;                       // [PROD_WARPS][LIST_BYTES]
  const uint32_t aseq = smem_u32(lists + PROD_WARPS * (PAIR ? 2 : 1) * LIST_BYTES);             // [MAX_A] u32: stage number + 1 in slot

  // warp index through a broadcast shuffle: the compiler then knows it is warp-uniform (role branches, barrier
  // addresses and slot numbers stay in uniform registers instead of per-lane copies with R2UR waterfalls)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // timeline marks: dbg[((role * 256 + stage) * 8 + event)] = clock64(), CTA 0 only, first 256 stages
  // (compiled in only with -DSCN_TC_TIMELINE: the marks cost the single-warp issue loops real time)
  auto mark = [&](int role, int stage, int ev) {
#ifdef SCN_TC_TIMELINE
    if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && stage < 256)
      p.dbg[((size_t)role * 256 + stage) * 8 + ev] = (unsigned long long)clock64();
#else
    (void)role; (void)stage; (void)ev;
#endif
  };

  if (warp == WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < SA; ++s) {
        st_release_u32(aseq + 4u * (uint32_t)s, 0u);
        mbar_init(afull(s), 32);                  // the 32 lanes of the owning producer warp, each when its copies landed
        mbar_init(aempty(s), 1);                  // one tcgen05.commit
      }
      for (int s = 0; s < SB; ++s) {
        mbar_init(bfull(s), 1);                   // the loader's expect_tx arrival (+ complete_tx bytes)
        mbar_init(bempty(s), p.NM);               // one tcgen05.commit per issuing warp
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(accf(b), p.NM);
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

  // This CTA's contiguous range of tiles [tile_lo, tile_hi): an even split of the tiles over the grid, cut into groups
  // of T tiles (only the last group of a CTA can be partial; its missing tiles are not processed by anybody).
  const int T = p.T;
  const int tile_lo = (int)(((int64_t)p.num_tiles * blockIdx.x) / gridDim.x);
  const int tile_hi = (int)(((int64_t)p.num_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int my_tiles = tile_hi - tile_lo;
  const int my_groups = (my_tiles + T - 1) / T;
  auto tiles_in_group = [&](int g) { return my_tiles - g * T < T ? my_tiles - g * T : T; };
  const int NSTEP = PAIR ? (p.K + 1) / 2 : p.K;            // offsets (offset pairs) per tile
  const int Q = NSTEP * NCH;                                // stages per tile

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // ================================ A producers (warp pw <-> A slot pw) ====================
    // NPW = SA/2 warps are active; warp pw owns A slots pw and pw + NPW and alternates between them (stage n -> warp
    // n mod NPW, slot n mod SA), so a slot has one producer and its mbarrier sees consecutive phases (parity-safe).
    // A single warp issues roughly one dependent instruction per 5 cycles, so the per-stage instruction count IS the
    // gather throughput: the list holds ready-made (source row, swizzled destination address) pairs, and an item
    // costs one LDS.64, one IMAD.WIDE, one LOP3 and one 16-byte LDGSTS per lane.  The copies are asynchronous: the
    // warp builds and issues its next stage (other slot) while this one is in flight, then waits, fences and publishes.
    const int pw = warp - EPI_WARPS;
    const int NPW = SA >> 1;
    if (pw < NPW) {
      constexpr int LPI = PAIR ? 4 : 8;                    // lanes per item (a 128-byte row, or a 64-byte half row)
      constexpr int IPP = 32 / LPI;                        // items per pass
      const int chunk = lane % LPI, sub = lane / LPI;
      int2* list = reinterpret_cast<int2*>(lists + pw * (PAIR ? 2 : 1) * LIST_BYTES);   // .x source row offset / 16 B (-1: zeros), .y smem address
      const uint32_t row_vec = (uint32_t)p.n_in >> 3;      // 16-byte units per feature row (list offsets are in these units)
      const int last_chunks = p.last_kc >> 3;
      const uint32_t lt = (1u << lane) - 1u;
      const uint32_t csw = (uint32_t)chunk << 4;
      const uint32_t slot_stride = (uint32_t)NPW * A_BYTES;
      const int total = my_tiles * Q;                      // stages of this CTA, in MMA order (group, stage, tile)
      // swizzled address of (row 32i + lane, 16-byte chunk 0) in slot pw (upper half, PAIR: ^ 0x40; other slot: + stride)
      uint32_t dlo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t r = 32u * i + lane;
        dlo[i] = a_base + (uint32_t)pw * A_BYTES + (r << 7) + ((r & 7u) << 4);
      }

      // stage cursor (tile t of stage q of group g)
      int t = 0, q = 0, g = 0;
      auto advance = [&](int& t_, int& q_, int& g_, int by) {
        t_ += by;
        int tv = tiles_in_group(g_);
        while (t_ >= tv && g_ < my_groups) {
          t_ -= tv;
          if (++q_ == Q) { q_ = 0; ++g_; tv = tiles_in_group(g_); }
        }
      };
      advance(t, q, g, pw);
      // neighbour indices of a stage: lane l holds rows l, 32+l, 64+l, 96+l of the tile (PAIR: of both offsets)
      auto load_idx = [&](int t_, int q_, int g_, int (&jl)[4], int (&jh)[4]) {
        const int step = PAIR ? q_ : q_ / NCH;
        const int64_t tile = (int64_t)tile_lo + (int64_t)g_ * T + t_;
        const int k0 = PAIR ? 2 * step : step;
        const int32_t* src = p.nbr + (int64_t)k0 * p.n_pad + tile * BM + lane;
        const bool hi_ok = PAIR && (k0 + 1 < p.K);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          jl[i] = ldg_nc32(src + 32 * i);
          jh[i] = hi_ok ? ldg_nc32(src + p.n_pad + 32 * i) : -1;
        }
      };

      int jl[4], jh[4];
      if (pw < total) load_idx(t, q, g, jl, jh);
      int it = 0;
      for (int n = pw; n < total; n += NPW, ++it) {
        const int slot = pw + (it & 1) * NPW;
        const uint32_t soff = (it & 1) ? slot_stride : 0u;
        mark(0, n, 0);
        const int cc = PAIR ? 0 : q % NCH;
        const bool full = (q == 0);                        // first stage of a tile: unmasked MMA, every row written
        // ---- live rows -> list (full stage: every row, missing ones as zeros) --------------------------------
        uint32_t B[4], H[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          B[i] = __ballot_sync(0xffffffffu, jl[i] >= 0);
          H[i] = PAIR ? __ballot_sync(0xffffffffu, jh[i] >= 0) : 0u;
        }
        int nlive = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pos = full ? 32 * i + lane : nlive + __popc(B[i] & lt);
          if (full || jl[i] >= 0) list[pos] = make_int2(jl[i] >= 0 ? (int)((uint32_t)jl[i] * row_vec) : -1, (int)(dlo[i] + soff));
          nlive += full ? 32 : __popc(B[i]);
        }
        if (PAIR) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int pos = full ? 128 + 32 * i + lane : nlive + __popc(H[i] & lt);
            if (full || jh[i] >= 0) list[pos] = make_int2(jh[i] >= 0 ? (int)((uint32_t)jh[i] * row_vec) : -1, (int)((dlo[i] + soff) ^ 0x40u));
            nlive += full ? 32 : __popc(H[i]);
          }
        }
        __syncwarp();
        uint32_t mword = 0;                                // lanes 0..7: the 8 words of the slot's lane masks
        if (lane < 8) {
          const uint32_t lo = (lane & 2) ? ((lane & 1) ? B[3] : B[2]) : ((lane & 1) ? B[1] : B[0]);
          const uint32_t hi = (lane & 2) ? ((lane & 1) ? H[3] : H[2]) : ((lane & 1) ? H[1] : H[0]);
          mword = ~(lane < 4 ? lo : hi);
        }
        mark(0, n, 1);
        // prefetch the next stage's indices (consumed in the next iteration)
        advance(t, q, g, NPW);
        if (n + NPW < total) load_idx(t, q, g, jl, jh);

        // the MMAs that read this slot's previous stage (two iterations ago) have retired
        mbar_wait(aempty(slot), (((uint32_t)it >> 1) & 1u) ^ 1u);
        if (lane < 8) amask[slot * 8 + lane] = mword;
        mark(0, n, 2);
        const int npass = (nlive + IPP - 1) / IPP;
        const bool lane_on = PAIR ? true : chunk < (cc == NCH - 1 ? last_chunks : 8);
        const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.in) + ((uint32_t)cc * 128u + csw);
        if (lane_on) {
          if (!full) {
            // hot path, ~6 instructions per item: LDS.64, IMAD.WIDE (source address), LOP3 (destination), LDGSTS
            for (int p0 = 0; p0 < npass; p0 += 6) {        // 6 list entries are read before their copies are issued
              int2 e[6];
#pragma unroll
              for (int u = 0; u < 6; ++u) {
                const int item = (p0 + u) * IPP + sub;
                e[u] = make_int2(0, 0);
                if (item < nlive) e[u] = list[item];
              }
#pragma unroll
              for (int u = 0; u < 6; ++u)
                if (e[u].y != 0) cp_async16((uint32_t)e[u].y ^ csw, src0 + ((uint64_t)(uint32_t)e[u].x << 4), 16u);
            }
          } else {
            // first stage of a tile: every row is written, missing neighbours as zeros (src-size 0 reads nothing)
            for (int pass = 0; pass < npass; ++pass) {
              const int2 e = list[pass * IPP + sub];
              const bool live = e.x != -1;
              cp_async16((uint32_t)e.y ^ csw, src0 + ((uint64_t)(live ? (uint32_t)e.x : 0u) << 4), live ? 16u : 0u);
            }
          }
        }
        // Landing is signalled by the copies themselves: every lane's arrival on afull(slot) fires when its cp.async
        // have completed.  The sequence flag only tells the issuing warps WHICH stage the slot now holds (and the
        // parity of the landing phase to wait for); the producer never waits for its own copies.
        cp_async_arrive_noinc(afull(slot));
        __syncwarp();                                      // all lanes have read the list (rewritten next iteration)
        if (lane == 0)
          st_release_u32(aseq + 4u * (uint32_t)slot, ((uint32_t)n + 1u) | ((((uint32_t)it >> 1) & 1u) << 31));
        mark(0, n, 3);
      }
    }
  } else if (warp == WARP_BLOAD) {
    // ================================ weight-tile loader (1 elected lane, bulk async copies) ==
    // The tiles (n_out x 128 B each, already in the swizzled shared-memory image) stream from L2 in stage order; the
    // copy engine signals bfull itself (complete_tx), so landing needs no warp and no proxy fence.
    const int total_b = my_groups * Q;
    int bslot = 0;
    uint32_t bround = 0;
    for (int i = 0; i < total_b; ++i) {
      mark(2, i, 0);
      mbar_wait(bempty(bslot), (bround & 1u) ^ 1u);
      mark(2, i, 1);
      if (elect_one()) {
        mbar_expect_tx(bfull(bslot), b_bytes);
        bulk_g2s(b_base + (uint32_t)bslot * b_bytes, p.bimg + (size_t)(i % Q) * b_bytes, b_bytes, bfull(bslot));
      }
      __syncwarp();
      if (++bslot == SB) { bslot = 0; ++bround; }
    }
  } else if (warp >= WARP_MMA && warp < WARP_MMA + MMA_WARPS) {
    // ================================ MMA issuers (warp-uniform loops, 1 elected lane issues) =
    // The per-stage issue path (two mbarrier waits, mask fetch, 4 UTCHMMA, commit) costs one warp several hundred
    // cycles, more than the tensor pipe needs for the stage, so the T tiles of a group (independent accumulators)
    // are dealt to NM = min(T, 4) issuing warps: warp m issues tiles m, m + NM, ...
    const int m = warp - WARP_MMA, NM = p.NM;
    if (m < NM) {
      const uint32_t idesc = make_idesc(p.n_out);
      const uint64_t da0 = make_desc_sw128(a_base), db0 = make_desc_sw128(b_base);
      const int nbuf = p.nbuf;
      int bslot = 0;
      uint32_t bround = 0;
      int aslot = m % SA;                                             // slot of this warp's next stage (n mod SA), kept incrementally
      uint32_t seq = (uint32_t)m + 1u;                                // its sequence number (n + 1)
      for (int g = 0; g < my_groups; ++g) {
        const int buf = nbuf == 2 ? (g & 1) : 0;
        const uint32_t use = (uint32_t)(nbuf == 2 ? (g >> 1) : g);    // how many times this buffer has been used before
        const int tv = tiles_in_group(g);
        mbar_wait(acce(buf), (use & 1u) ^ 1u);                        // epilogue has drained this accumulator buffer
        tc_fence_after();
        for (int q = 0; q < Q; ++q) {
          const int cc = PAIR ? 0 : q % NCH;
          const int step = PAIR ? q : q / NCH;
          mark(3, g * Q + q, 0);
          mbar_wait(bfull(bslot), bround & 1u);
          mark(3, g * Q + q, 1);
          const uint64_t db = db0 + (uint64_t)(((uint32_t)bslot * b_bytes) >> 4);
          // PAIR with an odd K: the last stage holds one offset only, its upper 32 channels are never written
          const int nk = (p.exp & 2) ? 0 : (PAIR ? ((2 * step + 1 < p.K) ? 4 : 2) : ((cc == NCH - 1 ? p.last_kc : KC) >> 4));
          int t = m;
          for (; t < tv; t += NM) {
            mark(1, (int)seq - 1, 0);
            {
              uint32_t spins = 0, f;
              while (((f = ld_acquire_u32(aseq + 4u * (uint32_t)aslot)) & 0x7fffffffu) != seq)   // slot holds stage n?
                if (++spins > SPIN_LIMIT) __trap();
              __syncwarp();
              mbar_wait(afull(aslot), f >> 31);              // ... and its rows have landed (phase parity from the flag)
            }
            mark(1, (int)seq - 1, 1);
            tc_fence_after();
            // disable-output-lane masks published by the stage's producer: bit r set <=> output row r has no
            // neighbour at this offset (its A row is stale).  Every lane loads the same words; the ballots make
            // them provably warp-uniform so they are moved to uniform registers once.
            uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0;
            if (q > 0) {
              const uint4 mw = *reinterpret_cast<const uint4*>(amask + aslot * 8);
              m0 = __ballot_sync(0xffffffffu, (mw.x >> lane) & 1u);
              m1 = __ballot_sync(0xffffffffu, (mw.y >> lane) & 1u);
              m2 = __ballot_sync(0xffffffffu, (mw.z >> lane) & 1u);
              m3 = __ballot_sync(0xffffffffu, (mw.w >> lane) & 1u);
              if (PAIR) {
                const uint4 hw = *reinterpret_cast<const uint4*>(amask + aslot * 8 + 4);
                h0 = __ballot_sync(0xffffffffu, (hw.x >> lane) & 1u);
                h1 = __ballot_sync(0xffffffffu, (hw.y >> lane) & 1u);
                h2 = __ballot_sync(0xffffffffu, (hw.z >> lane) & 1u);
                h3 = __ballot_sync(0xffffffffu, (hw.w >> lane) & 1u);
              }
            }
            const uint64_t da = da0 + (uint64_t)((uint32_t)aslot * (A_BYTES >> 4));
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u + (uint32_t)(t * p.n_out);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < nk) {
                  const bool hi = PAIR && kk >= 2;
                  umma_masked(tmem_d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (q > 0 || kk > 0) ? 1u : 0u,
                              hi ? h0 : m0, hi ? h1 : m1, hi ? h2 : m2, hi ? h3 : m3);
                }
              }
              umma_commit(aempty(aslot));                              // frees the A slot when these MMAs retire
            }
            __syncwarp();
            mark(1, (int)seq - 1, 2);
            aslot += NM; if (aslot >= SA) aslot -= SA;                 // NM <= 4 <= SA
            seq += (uint32_t)NM;
          }
          // this warp's next stage is tile m of the next q (or group): skip the tiles of other warps in between
          {
            const int adv = tv - t + m;                                // stages from (q, t) to (q + 1, m): tv - t + m, may be negative
            int a2 = aslot + adv;
            while (a2 < 0) a2 += SA;
            while (a2 >= SA) a2 -= SA;
            aslot = a2;
            seq = (uint32_t)((int)seq + adv);
          }
          if (elect_one()) umma_commit(bempty(bslot));
          __syncwarp();
          if (++bslot == SB) { bslot = 0; ++bround; }
        }
        if (elect_one()) umma_commit(accf(buf));                       // this warp's accumulators of the group are complete
        __syncwarp();
      }
    }
  } else if (warp < EPI_WARPS) {
    // ================================ epilogue (warps 0..3) ==================================
    for (int g = 0; g < my_groups; ++g) {
      const int buf = p.nbuf == 2 ? (g & 1) : 0;
      mbar_wait_sleep(accf(buf), (uint32_t)(p.nbuf == 2 ? (g >> 1) : g) & 1u);
      tc_fence_after();
      const int tv = tiles_in_group(g);
      for (int t = 0; t < tv; ++t) {
        const int64_t tile = (int64_t)tile_lo + (int64_t)g * T + t;
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

// Debug: device buffer of 4 roles x 256 stages x 8 marks (uint64) filled by CTA 0 of the following launches; NULL disables.
static unsigned long long* g_tc_dbg = nullptr;
extern "C" void scn_tc_debug_timeline(void* device_buffer) { g_tc_dbg = (unsigned long long*)device_buffer; }

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

// n_in == 32: two offsets share a 64-channel stage
static bool tc_pair(int n_in) { return n_in == 32; }

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

// conv_tcl.cu: the same convolution driven by precomputed stage lists (experimental, opt-in)
int scn_tcl_forward(const __nv_bfloat16* in, int64_t n_in_rows, const int32_t* nbr, int K, int64_t n_rows,
                    int64_t n_pad, int n_in, int n_out, const void* bimg, const float* bias, __nv_bfloat16* out,
                    const void* lists, unsigned long long* dbg, int exp_flags, cudaStream_t s);

// lists: stage lists of the table (stage_lists.cuh) or null; only a SUBMANIFOLD table (K odd, centre offset =
// identity, n_rows == n_in_rows) may come with lists, and the call is then routed to k_conv_tcl.
int scn_tc_forward(const __nv_bfloat16* in, int64_t n_in_rows, const int32_t* nbr, int K, int64_t n_rows,
                   int64_t n_pad, int n_in, int n_out, const void* bimg, const float* bias, __nv_bfloat16* out,
                   const void* lists, cudaStream_t s) {
  if ((uint64_t)n_in_rows * (uint64_t)(n_in >> 3) >= 0xffffffffull) return SCN_ERR_UNSUPPORTED;   // 32-bit offsets in 16-byte units (64 GB)
  tc::Params p;
  {
    static int exp_flags = -1;
    if (exp_flags < 0) {
      const char* e = std::getenv("SCN_B200_TC_EXP");
      exp_flags = e ? std::atoi(e) : 0;
    }
    p.exp = exp_flags;
  }
  if (lists != nullptr && (K & 1) == 1 && n_rows == n_in_rows && n_pad == (n_rows + tc::BM - 1) / tc::BM * tc::BM)
    return scn_tcl_forward(in, n_in_rows, nbr, K, n_rows, n_pad, n_in, n_out, bimg, bias, out, lists, g_tc_dbg, p.exp, s);
  p.dbg = g_tc_dbg;
  p.in = in; p.nbr = nbr; p.bimg = (const unsigned char*)bimg; p.bias = bias; p.out = out;
  p.n_rows = n_rows; p.n_pad = n_pad; p.K = K; p.n_in = n_in; p.n_out = n_out;
  const int nch = (n_in + tc::KC - 1) / tc::KC;
  p.last_kc = n_in - (nch - 1) * tc::KC;
  p.num_tiles = (int)((n_rows + tc::BM - 1) / tc::BM);
  // tiles per group: as many accumulators as fit in half of TMEM (weight-tile reuse, and one issuing warp per
  // tile up to MMA_WARPS); the tiles themselves are split evenly over the CTAs, so T does not unbalance the grid
  int T = 256 / n_out;
  if (T < 1) T = 1;
  if (T > 8) T = 8;
  p.nbuf = 2;
  int max_ctas = kNumSMs;
  {
    // developer knob: SCN_B200_TC_GRID=<CTAs> caps the grid (fewer CTAs, more tiles per group: less weight re-streaming)
    static int force_grid = -1;
    if (force_grid < 0) {
      const char* e = std::getenv("SCN_B200_TC_GRID");
      force_grid = e ? std::atoi(e) : 0;
    }
    if (force_grid > 0 && force_grid < max_ctas) max_ctas = force_grid;
  }
  const int per_cta = (p.num_tiles + max_ctas - 1) / max_ctas;
  {
    // One group in all 512 TMEM columns instead of two alternating halves: the epilogue then no longer overlaps the
    // next group's MMAs, but T doubles (more issuing warps, more weight-tile reuse).  Worth it when a CTA has a
    // single group anyway, or when half of TMEM holds one tile only.
    int t1 = 512 / n_out;
    if (t1 > 8) t1 = 8;
    if (t1 > T && (per_cta <= t1 || T == 1)) { T = t1; p.nbuf = 1; }
  }
  if (T > per_cta) T = per_cta;
  {
    // developer knob (sweeps of weight-tile re-streaming vs epilogue overlap): SCN_B200_TC_T=<tiles per group>
    static int force_t = -1;
    if (force_t < 0) {
      const char* e = std::getenv("SCN_B200_TC_T");
      force_t = e ? std::atoi(e) : 0;
    }
    if (force_t > 0) {
      T = force_t;
      if (T > 8) T = 8;
      if (T * n_out > 512) T = 512 / n_out;
      if (T > per_cta) T = per_cta;
      if (T < 1) T = 1;
      p.nbuf = T * n_out <= 256 ? 2 : 1;
    }
  }
  p.T = T;
  p.NM = T < tc::MMA_WARPS ? T : tc::MMA_WARPS;
  p.num_groups = 0;
  const uint32_t b_bytes = (uint32_t)n_out * 128u;
  // weight ring: up to 4 tiles (3 in flight behind the one being consumed), within ~72 KB
  {
    int sb = (int)((72u * 1024u) / b_bytes);
    if (sb > tc::MAX_B) sb = tc::MAX_B;
    if (sb < 2) sb = 2;
    p.SB = sb;
  }
  const bool pair = tc_pair(n_in);
  constexpr int NBAR = 2 * tc::MAX_A + 2 * tc::MAX_B + 4;
  const uint32_t fixed = 1024u + (uint32_t)p.SB * b_bytes + 8u * NBAR + 16u + (uint32_t)tc::MAX_A * tc::MASK_BYTES +
                         (uint32_t)tc::PROD_WARPS * tc::LIST_BYTES * (pair ? 2u : 1u) + 4u * tc::MAX_A + 12u;
  const uint32_t budget = 226u * 1024u;
  int SA = (int)((budget - fixed) / tc::A_BYTES);
  if (SA > tc::MAX_A) SA = tc::MAX_A;
  SA &= ~1;                                      // two A slots per producer warp
  if (SA < 4) return SCN_ERR_UNSUPPORTED;
  p.SA = SA;
  size_t smem = (size_t)fixed + (size_t)SA * tc::A_BYTES;
  int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  auto launch = [&](auto kern) -> int {
    SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, tc::THREADS, smem, s>>>(p);
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
