// k_conv_tcl: the tcgen05 gather-GEMM convolution of conv_tc.cu driven by precomputed STAGE LISTS (stage_lists.cuh).
//     out[o, :] = bias + sum_k  in[nbr[k][o], :] . B_k            (bf16 operands, fp32 accumulate)
//
// EXPERIMENTAL, OFF BY DEFAULT (SCN_B200_STAGE_LISTS=1 turns it on in the Python layer): written after the round's GPU
// budget was spent, compiled for sm_100a but not yet run on a GPU.  conv_tc.cu stays the validated product kernel and
// is untouched by this file; tests/test_gpu_stage_lists.py holds the parity cases for this one.
//
// Same CTA organisation, pipelines and synchronisation as k_conv_tc (read its header first): persistent CTA per SM,
// groups of T tiles in TMEM, 64-channel stages, bulk-copied weight tiles, LDGSTS row gather into SWIZZLE_128B A slots
// that signal their own landing, masked tcgen05.mma issued by up to 4 warps, 4 epilogue warps.  What changes is the
// producer warps' per-stage work, which bounds k_conv_tc at the shallow levels:
//   * k_conv_tc builds every stage from the neighbour table: 128 (PAIR: 256) index loads, 4-8 ballots, prefix sums,
//     list stores -- ~250 of the ~320 instructions a 32-channel stage costs its warp -- and every convolution that
//     shares the rulebook (8 layers x forward + dgrad per level) repeats it;
//   * here the list of a stage -- (source row, swizzled offset) per live row --, its length and its disable-lane mask
//     come from the stage-list buffer built once per rulebook: the warp fetches a 24-byte header two stages ahead,
//     starts ONE bulk async copy of the list (two in PAIR mode) one stage ahead into the other of its two list
//     buffers, and when the A slot is free copies the mask words and issues the LDGSTS straight from the landed list;
//   * the stage holding the CENTRE offset of the submanifold filter (every row is its own neighbour) is issued first
//     and unmasked, so it initialises the accumulator and no stage is ever written in full or zero-filled; the other
//     steps follow in ascending order (nat_step), and the weight loader fetches the image tiles in that order.
//     PAIR mode (32 channels, two offsets per stage): the centre shares its stage with a neighbouring offset; that
//     stage's MMAs start with the centre half (unmasked, accumulate = 0) and then issue the other half masked.
// HBM: the lists are 8 bytes per (in,out) pair plus 24 bytes per stage instead of 4*K bytes per row of the table
// (level 0 of the bench: ~54 instead of 108 bytes per row).
//
// Replaces SCN's dConvolution_KMxKN_forwardA/B; reference call sites
// src/networks/sparse_building_blocks.py:29-34; src/networks/resnet.py:30-36,44-50.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "stage_lists.cuh"
#include "tc_ptx.cuh"

namespace tcl {

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
  // stage lists of the table (stage_lists.cuh)
  const unsigned char* lists;
  int cstep;                    // step (offset, or offset pair in PAIR mode) that holds the centre offset: issued first
  int chalf;                    // PAIR: half of that stage's 64 channels the centre offset occupies (0 lower, 1 upper)
};

// NCH: 64-channel chunks per offset = ceil(n_in / 64).  PAIR (n_in == 32, NCH == 1): one stage holds TWO offsets,
// 32 channels each (chunks 0-3 from offset 2q, chunks 4-7 from offset 2q+1), halving the stage count.
template <int NCH, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) k_conv_tcl(const Params p) {
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
  constexpr int NBAR0 = 2 * MAX_A + 2 * MAX_B + 4;         // 38 -> 304 bytes (a multiple of 16: amask stays 16-byte aligned)
  auto lbar = [&](int w, int b) { return bar0 + 8u * (uint32_t)(NBAR0 + 2 * w + b); };   // list landed in buffer b of warp w
  constexpr int NBAR = NBAR0 + 2 * PROD_WARPS;                                            // 50 -> 400 bytes
  constexpr int LIST_WARP_BYTES = (PAIR ? 2 : 1) * LIST_BYTES * 2;                        // two list buffers per warp
  unsigned char* tail = gbase + (size_t)SA * A_BYTES + (size_t)SB * b_bytes + 8 * NBAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail);
  uint32_t* amask = reinterpret_cast<uint32_t*>(tail + 16);                    // [MAX_A][8], 16-byte aligned
  unsigned char* lists = tail + 16 + MAX_A * MASK_BYTES;                       // [PROD_WARPS][LIST_BYTES]
  const uint32_t aseq = smem_u32(lists + PROD_WARPS * LIST_WARP_BYTES);                         // [MAX_A] u32: stage number + 1 in slot

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
      for (int w = 0; w < PROD_WARPS; ++w)
        for (int b = 0; b < 2; ++b) mbar_init(lbar(w, b), 1);     // the issuing lane's expect_tx arrival (+ complete_tx bytes)
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
  // stage order: the centre step first, the others in ascending order
  auto nat_step = [&](int step) { return step == 0 ? p.cstep : (step <= p.cstep ? step - 1 : step); };

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // ================================ A producers (warp pw <-> A slot pw) ====================
    // NPW = SA/2 warps are active; warp pw owns A slots pw and pw + NPW and alternates between them (stage n -> warp
    // n mod NPW, slot n mod SA), so a slot has one producer and its mbarrier sees consecutive phases (parity-safe).
    // A single warp issues roughly one dependent instruction per 5 cycles, so the per-stage instruction count IS the
    // gather throughput: the landed list holds ready-made (source row, swizzled offset) pairs, and an item costs one
    // LDS.64, one IMAD.WIDE, one LOP3 + IADD and one 16-byte LDGSTS per lane.  Everything else about a stage -- its
    // header (two stages ahead) and its list (one stage ahead) -- is in flight while the previous stages are issued.
    const int pw = warp - EPI_WARPS;
    const int NPW = SA >> 1;
    if (pw < NPW) {
      constexpr int LPI = PAIR ? 4 : 8;                    // lanes per item (a 128-byte row, or a 64-byte half row)
      constexpr int IPP = 32 / LPI;                        // items per pass
      const int chunk = lane % LPI, sub = lane / LPI;
      int2* list = reinterpret_cast<int2*>(lists + pw * LIST_WARP_BYTES);   // this warp's two list buffers (.x source row, .y offset in the slot)
      const int last_chunks = p.last_kc >> 3;
      const uint32_t csw = (uint32_t)chunk << 4;
      const int total = my_tiles * Q;                      // stages of this CTA, in MMA order (group, stage, tile)

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
      {
        // ---------------- precomputed stage lists ----------------
        const int2* hdr = reinterpret_cast<const int2*>(p.lists + sl::hdr_offset());
        const uint32_t* msk = reinterpret_cast<const uint32_t*>(p.lists + sl::msk_offset(p.num_tiles, p.K));
        const unsigned char* ent = p.lists + sl::ent_offset(p.num_tiles, p.K);
        if (pw == 0 && lane == 0) {                        // the buffer was built for this table shape
          const uint32_t* head = reinterpret_cast<const uint32_t*>(p.lists);
          if (head[1] != (uint32_t)p.K || head[2] != (uint32_t)p.num_tiles || head[3] != sl::MAGIC) __trap();
        }
        const uint32_t row_bytes = (uint32_t)p.n_in * 2u;
        const uint32_t lbuf_u32 = smem_u32(list);          // buffer b: + b * LIST_WARP_BYTES / 2; PAIR second list: + LIST_BYTES
        constexpr uint32_t LBUF = LIST_WARP_BYTES / 2;
        struct Hd { int off0, cnt0, off1, cnt1, cc; uint32_t mword; };
        // header of the stage under the cursor: where its list(s) are, how long, its mask words (lanes 0..7), its chunk
        auto load_hd = [&](int t_, int q_, int g_) {
          Hd h;
          const int step = PAIR ? q_ : q_ / NCH;
          const int ns = nat_step(step);
          const int64_t tile = (int64_t)tile_lo + (int64_t)g_ * T + t_;
          const int k0 = PAIR ? 2 * ns : ns;
          const bool has1 = PAIR && (k0 + 1 < p.K);
          const size_t s0 = (size_t)tile * (size_t)p.K + (size_t)k0;
          const int2 a = __ldg(hdr + s0);
          int2 b = make_int2(0, 0);
          if (has1) b = __ldg(hdr + s0 + 1);
          h.off0 = a.x; h.cnt0 = a.y; h.off1 = b.x; h.cnt1 = b.y;
          h.cc = PAIR ? 0 : q_ % NCH;
          h.mword = 0xffffffffu;
          if (lane < 4) h.mword = __ldg(msk + s0 * 4 + lane);
          else if (lane < 8 && has1) h.mword = __ldg(msk + (s0 + 1) * 4 + (lane - 4));
          return h;
        };
        // one bulk copy per list into buffer b of this warp; the copy engine signals lbar itself
        auto issue_list = [&](int b, const Hd& h) {
          const uint32_t bytes0 = (uint32_t)h.cnt0 * 8u, bytes1 = PAIR ? (uint32_t)h.cnt1 * 8u : 0u;   // counts are multiples of 8 entries
          if (elect_one()) {
            mbar_expect_tx(lbar(pw, b), bytes0 + bytes1);
            const uint32_t dst = lbuf_u32 + (uint32_t)b * LBUF;
            if (bytes0) bulk_g2s(dst, ent + (size_t)(uint32_t)h.off0 * 8u, bytes0, lbar(pw, b));
            if (bytes1) bulk_g2s(dst + LIST_BYTES, ent + (size_t)(uint32_t)h.off1 * 8u, bytes1, lbar(pw, b));
          }
          __syncwarp();
        };
        Hd hA = {0, 0, 0, 0, 0, 0xffffffffu}, hB = hA;
        if (pw < total) {
          hA = load_hd(t, q, g);
          issue_list(0, hA);
        }
        advance(t, q, g, NPW);
        if (pw + NPW < total) hB = load_hd(t, q, g);
        int it = 0;
        for (int n = pw; n < total; n += NPW, ++it) {
          const int slot = pw + (it & 1) * NPW;
          const uint32_t slot_base = a_base + (uint32_t)slot * A_BYTES;
          mark(0, n, 0);
          // the next stage's list -> the other buffer (last read two iterations ago, before that iteration's __syncwarp)
          if (n + NPW < total) issue_list((it + 1) & 1, hB);
          // header of the stage after that (in flight while this stage is issued)
          Hd hC = hB;
          advance(t, q, g, NPW);
          if (n + 2 * NPW < total) hC = load_hd(t, q, g);
          mark(0, n, 1);
          // the MMAs that read this slot's previous stage (two iterations ago) have retired
          mbar_wait(aempty(slot), (((uint32_t)it >> 1) & 1u) ^ 1u);
          if (lane < 8) amask[slot * 8 + lane] = hA.mword;
          // this stage's list has landed
          mbar_wait(lbar(pw, it & 1), ((uint32_t)it >> 1) & 1u);
          mark(0, n, 2);
          const bool lane_on = PAIR ? true : chunk < (hA.cc == NCH - 1 ? last_chunks : 8);
          const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.in) + ((uint32_t)hA.cc * 128u + csw);
          // keep the whole base in one register pair: otherwise the compiler leaves p.in in a uniform register and
          // adds it to every item's address (IADD3 + IADD3.X per item on top of the IMAD.WIDE)
          asm volatile("" : "+l"(src0));
          if (lane_on) {
            // Lists are whole groups of 8 entries (padded by repeating the last one), so there is no per-item predicate:
            // an item costs LDS.64, IMAD.WIDE (source), LOP3 + IADD (destination) and the 16-byte LDGSTS.
            auto copy = [&](const int2 e, uint32_t cx) {
              cp_async16(slot_base + ((uint32_t)e.y ^ cx), src0 + (uint64_t)(uint32_t)e.x * (uint64_t)row_bytes, 16u);
            };
#pragma unroll
            for (int half = 0; half < (PAIR ? 2 : 1); ++half) {
              const int cnt = half ? hA.cnt1 : hA.cnt0;
              const int2* L = reinterpret_cast<const int2*>(reinterpret_cast<const unsigned char*>(list) +
                                                            (size_t)(it & 1) * LBUF + (size_t)half * LIST_BYTES) + sub;
              const uint32_t cx = half ? (csw ^ 0x40u) : csw;
              if constexpr (PAIR) {                        // 8 items per pass: 16 entries per iteration, one pass may remain
                int i = 0;
                for (; i + 8 < cnt; i += 16) {
                  const int2 e0 = L[i], e1 = L[i + 8];
                  copy(e0, cx);
                  copy(e1, cx);
                }
                if (i < cnt) copy(L[i], cx);
              } else {                                     // 4 items per pass: 8 entries = two passes per iteration
                for (int i = 0; i < cnt; i += 8) {
                  const int2 e0 = L[i], e1 = L[i + 4];
                  copy(e0, cx);
                  copy(e1, cx);
                }
              }
            }
          }
          cp_async_arrive_noinc(afull(slot));
          __syncwarp();                                    // all lanes have read the list buffer (refilled next iteration)
          if (lane == 0)
            st_release_u32(aseq + 4u * (uint32_t)slot, ((uint32_t)n + 1u) | ((((uint32_t)it >> 1) & 1u) << 31));
          mark(0, n, 3);
          hA = hB;
          hB = hC;
        }
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
        int bq = i % Q;                                  // image tile of the stage (the image is in natural step order)
        bq = PAIR ? nat_step(bq) : nat_step(bq / NCH) * NCH + bq % NCH;
        bulk_g2s(b_base + (uint32_t)bslot * b_bytes, p.bimg + (size_t)bq * b_bytes, b_bytes, bfull(bslot));
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
          const int nk = (p.exp & 2) ? 0 : (PAIR ? ((2 * nat_step(step) + 1 < p.K) ? 4 : 2) : ((cc == NCH - 1 ? p.last_kc : KC) >> 4));
          // PAIR, first stage: the half that holds the centre offset is unmasked and issued first (it
          // initialises the accumulator), the other half is an ordinary masked offset
          const bool first_pair = PAIR && q == 0;
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
            if (q > 0 || first_pair) {
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
            if (first_pair) {
              if (p.chalf) { h0 = 0; h1 = 0; h2 = 0; h3 = 0; }
              else { m0 = 0; m1 = 0; m2 = 0; m3 = 0; }
            }
            const uint64_t da = da0 + (uint64_t)((uint32_t)aslot * (A_BYTES >> 4));
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u + (uint32_t)(t * p.n_out);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < nk) {
                  const int kx = (first_pair && p.chalf) ? (kk ^ 2) : kk;      // K16 slice of the stage this MMA reads
                  const bool hi = PAIR && kx >= 2;
                  umma_masked(tmem_d, da + (uint64_t)(kx * 2), db + (uint64_t)(kx * 2), idesc, (q > 0 || kk > 0) ? 1u : 0u,
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

}  // namespace tcl

static bool tcl_pair(int n_in) { return n_in == 32; }

// Launched by scn_tc_forward (conv_tc.cu) when the caller passed the stage lists of a submanifold table
// (K odd, centre offset = identity, n_rows == n_in_rows).  Same tile / ring sizing as k_conv_tc.
int scn_tcl_forward(const __nv_bfloat16* in, int64_t n_in_rows, const int32_t* nbr, int K, int64_t n_rows,
                    int64_t n_pad, int n_in, int n_out, const void* bimg, const float* bias, __nv_bfloat16* out,
                    const void* lists, unsigned long long* dbg, int exp_flags, cudaStream_t s) {
  if (!lists || (K & 1) == 0 || n_rows != n_in_rows) return SCN_ERR_ARG;
  if ((uint64_t)n_in_rows * (uint64_t)n_in * 2ull >= (1ull << 40)) return SCN_ERR_UNSUPPORTED;
  tcl::Params p;
  p.exp = exp_flags;
  p.dbg = dbg;
  p.lists = (const unsigned char*)lists;
  p.in = in; p.nbr = nbr; p.bimg = (const unsigned char*)bimg; p.bias = bias; p.out = out;
  p.n_rows = n_rows; p.n_pad = n_pad; p.K = K; p.n_in = n_in; p.n_out = n_out;
  const int nch = (n_in + tcl::KC - 1) / tcl::KC;
  p.last_kc = n_in - (nch - 1) * tcl::KC;
  p.num_tiles = (int)((n_rows + tcl::BM - 1) / tcl::BM);
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
  p.NM = T < tcl::MMA_WARPS ? T : tcl::MMA_WARPS;
  p.num_groups = 0;
  const uint32_t b_bytes = (uint32_t)n_out * 128u;
  {
    int sb = (int)((72u * 1024u) / b_bytes);
    if (sb > tcl::MAX_B) sb = tcl::MAX_B;
    if (sb < 2) sb = 2;
    p.SB = sb;
  }
  const bool pair = tcl_pair(n_in);
  const int c = (K - 1) / 2;                     // centre offset of the (odd) filter
  p.cstep = pair ? (c >> 1) : c;
  p.chalf = pair ? (c & 1) : 0;
  // mirrors the kernel's carve-up (two mbarriers and two list buffers per producer warp more than k_conv_tc)
  constexpr int NBAR = 2 * tcl::MAX_A + 2 * tcl::MAX_B + 4 + 2 * tcl::PROD_WARPS;
  const uint32_t fixed = 1024u + (uint32_t)p.SB * b_bytes + 8u * NBAR + 16u + (uint32_t)tcl::MAX_A * tcl::MASK_BYTES +
                         (uint32_t)tcl::PROD_WARPS * tcl::LIST_BYTES * (pair ? 2u : 1u) * 2u + 4u * tcl::MAX_A + 12u;
  const uint32_t budget = 226u * 1024u;
  int SA = (int)((budget - fixed) / tcl::A_BYTES);
  if (SA > tcl::MAX_A) SA = tcl::MAX_A;
  SA &= ~1;                                      // two A slots per producer warp
  if (SA < 4) return SCN_ERR_UNSUPPORTED;
  p.SA = SA;
  size_t smem = (size_t)fixed + (size_t)SA * tcl::A_BYTES;
  int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  auto launch = [&](auto kern) -> int {
    SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, tcl::THREADS, smem, s>>>(p);
    SCN_LAUNCH_CHECK();
    return SCN_OK;
  };
  if (pair) return launch(tcl::k_conv_tcl<1, true>);
  switch (nch) {
    case 1: return launch(tcl::k_conv_tcl<1, false>);
    case 2: return launch(tcl::k_conv_tcl<2, false>);
    case 3: return launch(tcl::k_conv_tcl<3, false>);
    case 4: return launch(tcl::k_conv_tcl<4, false>);
    default: return SCN_ERR_UNSUPPORTED;
  }
}
