// Fused ICL backward (model/SNAG_loss.py:98-126, gradient of the in-batch contrastive loss w.r.t. the normalised rows)
// for contraction widths Dpad <= 320 — the per-modality calls of a step (D = 300): 8 of the 10 icl_loss calls at the
// reference's own configuration, 12 of 14 with surface features.
//
//   dZ_i = sum_j G_ij y_j ,   G_ij = dL/dlogit_ij = ((cr_i + cc_j) E_ij - [i == j, cross part] dg_i) / tau ,
//   E_ij = exp(s_ij / tau - 1 / tau) ,  s_ij = z_i . y_j                               (Appendix A of SURVEY.md)
//
// The two-kernel form (sim_kernel<EpiIclBwd> writes G [Bp, 2Bp] in bf16, a split-K GEMM contracts it with the stacked
// embeddings) moves 2 x 1 GB per side and call at B = 16 384 and keeps the tensor pipe 40 % busy (ncu, profiles/r01c):
// at D = 300 a logits tile is 5 k-blocks of MMA against 32 K exponentials. Here G never leaves the SM
// (flash-attention-backward shape):
//
//   per unit = (problem, side, block of 128 anchors, column split):
//     X  [128 x Dpad]  anchors, shared memory, loaded once per unit (TMA, K-major, SWIZZLE_128B)
//     for every 64-column tile t of [other side ; this side]:
//       Y_t [64 x Dpad]           TMA -> 3-deep ring
//       MMA1  S  = X . Y_t^T      tcgen05.mma SS, 128x64xDpad, fp32 accumulator in TMEM (2 stages)
//       epi   P  = bf16(G(S))     tcgen05.ld -> exp2 / coefficients -> tcgen05.st: P stays in TMEM (2 buffers)
//       MMA2  dZ += P . Y_t       tcgen05.mma TS: A = P from TMEM, B = the SAME shared-memory tile read MN-major
//                                 (the 128-byte-swizzled rows TMA wrote are exactly the canonical MN-major SW128 atoms:
//                                 64 d-elements x 8 rows j, LBO = k-block stride, SBO = 1024 B)
//     dZ [128 x Dpad] fp32 accumulator (TMEM, 320 columns) -> global, one partial per column split
//
// TMEM: S 2 x 64 + P 2 x 32 + dZ 320 = 512 columns. Shared memory: X 80 KB + Y 3 x 40 KB (a 128-column tile would halve
// the number of MMA1 instructions, but two of its Y tiles and X do not fit in 227 KB).
// Roles: TMA producer warp; TWO issuing warps — one for MMA1, running ahead as far as the S stages and the Y ring
// allow, one for MMA2, following the epilogue — so that neither product waits behind the other's barriers in a single
// instruction stream; and two PAIRS of epilogue warpgroups that take alternate tiles (pair g owns S stage g and P
// buffer g), so that two tiles are always inside the TMEM-read -> exp2 -> pack -> TMEM-write chain.
// Several problems (the calls of one step share B) are batched into one launch so that B = 3500 fills the machine.
#include <mutex>
#include "common.cuh"
#include "snag_internal.h"

namespace snag {

constexpr int FB_BM = 128;                     // anchors per unit (UMMA M)
constexpr int FB_BN = 64;                      // columns per tile (UMMA N of MMA1, K of MMA2)
constexpr int FB_BK = 64;                      // bf16 per k-block = one 128-byte swizzle row
constexpr int FB_MAX_KB = 5;                   // Dpad <= 320
constexpr int FB_YBUFS = 3;
constexpr int FB_XKB_BYTES = FB_BM * FB_BK * 2;              // 16 KB per k-block of X
constexpr int FB_YKB_BYTES = FB_BN * FB_BK * 2;              // 8 KB per k-block of a Y tile
constexpr int FB_X_BYTES = FB_MAX_KB * FB_XKB_BYTES;         // 80 KB
constexpr int FB_Y_BYTES = FB_MAX_KB * FB_YKB_BYTES;         // 40 KB per ring slot
constexpr int FB_BAR_BYTES = 256;
constexpr int FB_PAIR_THREADS = 256;           // a PAIR of warpgroups consumes one tile: WG h of the pair owns columns [32 h, 32 h + 32)
constexpr int FB_EPI_PAIRS = 2;                // pair g takes every other tile (those that land in S stage g / P buffer g), so two
                                               // tiles are in the epilogue at any time: its latency chain (TMEM read -> exp2 -> pack ->
                                               // TMEM write, ~2 800 clocks per tile with one pair — ncu: tensor pipe 52 % busy, issue
                                               // slots 26 %) is overlapped with itself instead of pacing the MMAs
constexpr int FB_EPI_THREADS = FB_EPI_PAIRS * FB_PAIR_THREADS;
constexpr int FB_THREADS = FB_EPI_THREADS + 96;
constexpr int FB_COEF_FLOATS = FB_EPI_PAIRS * 2 * FB_BN;     // per pair: the tile's column coefficients, double-buffered
constexpr int FB_SMEM_BYTES = 1024 + FB_X_BYTES + FB_YBUFS * FB_Y_BYTES + FB_BAR_BYTES + FB_COEF_FLOATS * 4;
constexpr int FB_TMEM_S = 0;                   // 2 stages x 64 fp32 columns
constexpr int FB_TMEM_P = 128;                 // 2 buffers x 32 columns (64 bf16 per lane)
constexpr int FB_TMEM_DZ = 192;                // 320 fp32 columns
constexpr int FB_MAX_PROB = 16;

struct alignas(64) FbProblem {
  CUtensorMap tm;          // stacked operand [a ; b ; a] of the call: [3 Bp, Dpad] bf16, box [64 rows x 64], SWIZZLE_128B
  const float* cr_a;       // [B] g_a[i] * exp(1/tau - lse_a[i])
  const float* cr_b;       // [B] the same for side b
  const float* dg;         // [B] g_a[i] + g_b[i]
  float* dz_a;             // [nsplit][128 * row_blocks][Dpad] partial dL/d(normalised a rows) of the launch's anchors
  float* dz_b;
  long long pad_[3];
};
static_assert(sizeof(FbProblem) == 192, "FbProblem layout");

struct FbParams {
  int n_prob, B, Bp, kblocks;
  int rb0;                 // first block of 128 anchors this launch covers (anchor sharding over ranks)
  int row_blocks;          // blocks of 128 anchors per side in this launch
  int nsplit;              // column splits per (problem, side, row block)
  int tiles_per_split;     // 64-column tiles per split (the last split may be shorter)
  int n_tiles;             // 2 Bp / 64
  int n_units;
  float scale_log2;        // log2(e) / tau
  float inv_tau;
  long long part_stride;   // floats between the partial outputs of consecutive splits
  FbProblem prob[FB_MAX_PROB];
};

// instruction descriptor as make_idesc_bf16, with B read MN-major ([16] b_major = 1)
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(int m, int n) { return make_idesc_bf16(m, n) | (1u << 16); }

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x (16 bf16 = 8 packed 32-bit columns) at a_tmem
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  const uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;       // disable-output-lane mask: none
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
}

#define FB_TMEM_ST16(taddr, w)                                                                                       \
  asm volatile(                                                                                                      \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                                \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"                                     \
      ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), \
        "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])                            \
      : "memory")
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct FbUnit {
  int prob, side, rb, split, t0, t1;
};
__device__ __forceinline__ FbUnit fb_decode(const FbParams& p, int u) {
  FbUnit q;
  q.split = u % p.nsplit;
  int v = u / p.nsplit;
  q.rb = v % p.row_blocks;
  v /= p.row_blocks;
  q.side = v & 1;
  q.prob = v >> 1;
  q.t0 = q.split * p.tiles_per_split;
  q.t1 = min(q.t0 + p.tiles_per_split, p.n_tiles);
  return q;
}
// a 64-column tile holds at least one valid column (tiles that lie entirely in the zero padding of a part are skipped
// by every role alike)
__device__ __forceinline__ bool fb_tile_valid(const FbParams& p, int t) {
  const int col0 = t * FB_BN;
  const int idx0 = col0 >= p.Bp ? col0 - p.Bp : col0;
  return idx0 < p.B;
}

// The valid tiles of a part are its first ceil(B / 64); of a unit's tiles [t0, t1) the valid ones are therefore two
// contiguous runs (one per part). FbValid enumerates them: n valid tiles, the j-th being tile(j).
struct FbValid {
  int lo0, n0, lo1, n;
  __device__ __forceinline__ int tile(int j) const { return j < n0 ? lo0 + j : lo1 + (j - n0); }
};
__device__ __forceinline__ FbValid fb_valid_tiles(const FbParams& p, int t0, int t1) {
  const int tpp = p.Bp / FB_BN;                              // tiles per part
  const int nvp = (p.B + FB_BN - 1) / FB_BN;                 // valid tiles per part
  FbValid v;
  v.lo0 = t0;
  v.n0 = max(0, min(t1, nvp) - v.lo0);
  v.lo1 = max(t0, tpp);
  v.n = v.n0 + max(0, min(t1, tpp + nvp) - v.lo1);
  return v;
}

__global__ void __launch_bounds__(FB_THREADS, 1) icl_bwd_fused_kernel(const __grid_constant__ FbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t x_base = base;
  const uint32_t y_base = base + FB_X_BYTES;
  const uint32_t bar0 = y_base + FB_YBUFS * FB_Y_BYTES;
  // barriers: 0 x_full, 1 x_empty, 2..4 y_full, 5..7 y_empty, 8..9 s_full, 10..11 s_empty, 12..13 p_full,
  //           14..15 p_empty, 16 dz_full, 17 dz_empty; then the TMEM base address slot
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_slot = bar0 + 8u * 18;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + FB_X_BYTES + FB_YBUFS * FB_Y_BYTES + 8 * 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cwarp = warp - FB_EPI_THREADS / 32;       // 0 = TMA producer, 1 = UMMA issuer of S, 2 = TMEM allocator + UMMA issuer of dZ; < 0: epilogue

  if (cwarp == 1 && lane == 0) {
    mbar_init(bar(0), 1);
    mbar_init(bar(1), 1);
    for (int i = 0; i < FB_YBUFS; ++i) { mbar_init(bar(2 + i), 1); mbar_init(bar(5 + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(8 + i), 1);
      mbar_init(bar(10 + i), FB_PAIR_THREADS);            // S stage i / P buffer i belong to epilogue pair i
      mbar_init(bar(12 + i), FB_PAIR_THREADS);
      mbar_init(bar(14 + i), 1);
    }
    mbar_init(bar(16), 1);
    mbar_init(bar(17), FB_EPI_THREADS);
    fence_mbar_init();
  }
  if (cwarp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int Dpad = p.kblocks * FB_BK;

  if (cwarp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const uint32_t leader = elect_one_sync();
    uint32_t xph = 0, yb = 0, yph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const FbUnit q = fb_decode(p, u);
      const CUtensorMap* tm = &p.prob[q.prob].tm;
      const int xrow0 = q.side * p.Bp + (p.rb0 + q.rb) * FB_BM;   // side a anchors: rows [0, Bp); side b: [Bp, 2 Bp)
      const int yrow0 = (1 - q.side) * p.Bp;                 // side a sweeps [b ; a] = rows [Bp, 3 Bp); side b [a ; b] = [0, 2 Bp)
      bool any = false;
      for (int t = q.t0; t < q.t1; ++t) any |= fb_tile_valid(p, t);
      if (!any) continue;
      mbar_wait(bar(1), xph ^ 1);                            // the previous unit's MMA1s no longer read X
      if (leader) {
        mbar_expect_tx(bar(0), static_cast<uint32_t>(p.kblocks) * FB_XKB_BYTES);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tma_load_2d(x_base + kb * FB_XKB_BYTES, tm, bar(0), kb * FB_BK, xrow0);
          tma_load_2d(x_base + kb * FB_XKB_BYTES + FB_XKB_BYTES / 2, tm, bar(0), kb * FB_BK, xrow0 + 64);
        }
      }
      __syncwarp();
      xph ^= 1;
      for (int t = q.t0; t < q.t1; ++t) {
        if (!fb_tile_valid(p, t)) continue;
        mbar_wait(bar(5 + yb), yph ^ 1);
        if (leader) {
          mbar_expect_tx(bar(2 + yb), static_cast<uint32_t>(p.kblocks) * FB_YKB_BYTES);
          for (int kb = 0; kb < p.kblocks; ++kb)
            tma_load_2d(y_base + yb * FB_Y_BYTES + kb * FB_YKB_BYTES, tm, bar(2 + yb), kb * FB_BK, yrow0 + t * FB_BN);
        }
        __syncwarp();
        if (++yb == FB_YBUFS) { yb = 0; yph ^= 1; }
      }
    }
  } else if (cwarp == 1) {
    // ------------------------------------------------------------------ UMMA issuer 1: S_t = X . Y_t^T
    // Two issuing warps: a 128x64x16 MMA occupies the tensor pipe for only 32-64 clocks, so 28 of them per tile plus
    // the barrier waits of BOTH products were more than one thread could issue in a tile's time. This warp runs ahead
    // as far as the S stages and the Y ring allow; the second one (below) follows the epilogue.
    const uint32_t leader = elect_one_sync();
    const uint32_t idesc1 = make_idesc_bf16(FB_BM, FB_BN);
    uint32_t xph = 0, yb = 0, yph = 0, sb = 0, sph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const FbUnit q = fb_decode(p, u);
      const FbValid vt = fb_valid_tiles(p, q.t0, q.t1);
      if (vt.n == 0) continue;
      mbar_wait(bar(0), xph);                                // X landed
      xph ^= 1;
      tc_fence_after();
      for (int j = 0; j < vt.n; ++j) {
        mbar_wait(bar(2 + yb), yph);                         // Y_t landed
        mbar_wait(bar(10 + sb), sph ^ 1);                    // accumulator stage drained by the epilogue
        tc_fence_after();
        if (leader) {
          const uint32_t tmem_s = tmem_base + FB_TMEM_S + sb * FB_BN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const uint64_t adesc = make_sdesc_k128(x_base + kb * FB_XKB_BYTES);
            const uint64_t bdesc = make_sdesc_k128(y_base + yb * FB_Y_BYTES + kb * FB_YKB_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16_ss(tmem_s, adesc + 2u * k4, bdesc + 2u * k4, idesc1, (kb | k4) != 0 ? 1u : 0u);
          }
          umma_commit(bar(8 + sb));                           // S_t complete -> epilogue
        }
        __syncwarp();
        if (++yb == FB_YBUFS) { yb = 0; yph ^= 1; }
        if (++sb == 2) { sb = 0; sph ^= 1; }
      }
      if (leader) umma_commit(bar(1));                        // every MMA1 of the unit has read X
      __syncwarp();
    }
  } else if (cwarp == 2) {
    // ------------------------------------------------------------------ UMMA issuer 2: dZ += P_t . Y_t
    const uint32_t leader = elect_one_sync();
    // dZ is up to 320 columns wide, a UMMA at most 256: two instructions of 192 + 128 rather than 256 + 64 — a 64-wide
    // instruction keeps the tensor pipe busy for 32 clocks but occupies it for about twice that
    const int n_lo = Dpad <= 256 ? Dpad : 192;
    const int n_hi = Dpad - n_lo;                            // 0 or 128
    const uint32_t idesc2_lo = make_idesc_bf16_bmn(FB_BM, n_lo);
    const uint32_t idesc2_hi = make_idesc_bf16_bmn(FB_BM, n_hi > 0 ? n_hi : 16);
    uint32_t yb = 0, yph = 0, pb = 0, pph = 0, dzph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const FbUnit q = fb_decode(p, u);
      const FbValid vt = fb_valid_tiles(p, q.t0, q.t1);
      if (vt.n == 0) continue;
      for (int j = 0; j < vt.n; ++j) {
        mbar_wait(bar(12 + pb), pph);                        // P_t is in TMEM (hence S_t was complete and Y_t had landed)
        mbar_wait(bar(2 + yb), yph);                         // this thread's own view of the Y_t barrier (already complete)
        if (j == 0) mbar_wait(bar(17), dzph ^ 1);            // the epilogue has read the previous unit's dZ
        tc_fence_after();
        if (leader) {
          const uint32_t ytile = y_base + yb * FB_Y_BYTES;
#pragma unroll
          for (int ks = 0; ks < FB_BN / 16; ++ks) {
            const uint32_t a_tmem = tmem_base + FB_TMEM_P + pb * 32 + ks * 8;
            // 16 rows j of the tile = 2 groups of 8 rows = 2048 B further along K
            const uint64_t bdesc = make_sdesc_mn128(ytile + ks * 2048, FB_YKB_BYTES);
            const uint32_t acc = (j == 0 && ks == 0) ? 0u : 1u;
            umma_bf16_ts(tmem_base + FB_TMEM_DZ, a_tmem, bdesc, idesc2_lo, acc);
            if (n_hi > 0)
              umma_bf16_ts(tmem_base + FB_TMEM_DZ + n_lo, a_tmem,
                           bdesc + static_cast<uint64_t>(((n_lo / FB_BK) * FB_YKB_BYTES) >> 4), idesc2_hi, acc);
          }
          umma_commit(bar(5 + yb));                           // Y slot free
          umma_commit(bar(14 + pb));                          // P buffer free
        }
        __syncwarp();
        if (++yb == FB_YBUFS) { yb = 0; yph ^= 1; }
        if (++pb == 2) { pb = 0; pph ^= 1; }
      }
      if (leader) umma_commit(bar(16));                       // dZ complete -> epilogue
      __syncwarp();
      dzph ^= 1;
    }
  } else if (cwarp < 0) {
    // ------------------------------------------------------------------ epilogue: two pairs of warpgroups
    const int tid = threadIdx.x;
    const int et = tid & 127;                                // row within the block == TMEM lane
    const int wg = tid >> 7;
    const int g = wg >> 1;                                   // pair: owns S stage g, P buffer g, i.e. every other valid tile
    const int h = wg & 1;                                    // which 32 of the tile's 64 columns
    const int pt = tid & (FB_PAIR_THREADS - 1);              // thread within the pair
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float nb = -p.scale_log2;
    float* coef = reinterpret_cast<float*>(smem + FB_X_BYTES + FB_YBUFS * FB_Y_BYTES + FB_BAR_BYTES) + g * 2 * FB_BN;
    uint32_t kglob = 0;                                      // valid tiles of this CTA's earlier units (stage = k & 1)
    uint32_t ph = 0;                                         // phase of this pair's S stage / P buffer
    uint32_t dzph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const FbUnit q = fb_decode(p, u);
      const FbProblem& pr = p.prob[q.prob];
      const float* cr_this = q.side ? pr.cr_b : pr.cr_a;
      const float* cr_other = q.side ? pr.cr_a : pr.cr_b;
      float* dz = (q.side ? pr.dz_b : pr.dz_a) + q.split * p.part_stride;
      const int gr0 = (p.rb0 + q.rb) * FB_BM;
      const int gr = gr0 + et;                               // batch index of this thread's anchor
      const long long orow_idx = static_cast<long long>(q.rb) * FB_BM + et;   // row of the (launch-local) output
      float* orow = dz + orow_idx * Dpad;
      const FbValid vt = fb_valid_tiles(p, q.t0, q.t1);
      if (vt.n == 0) {
        // a split that lies entirely in the padding: its partial gradient is zero
        for (int c = wg * 32; c < Dpad; c += 32 * (FB_EPI_THREADS / 128))
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) st_global_256(orow + c + 8 * hh, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0));
        continue;
      }
      const bool ok = gr < p.B;
      const float cr = ok ? __ldg(cr_this + gr) * p.inv_tau : 0.f;
      const float dg = ok ? __ldg(pr.dg + gr) * p.inv_tau : 0.f;
      // column coefficient of tile t's column pt (pt < 64), 1/tau folded in; zero in the padding
      auto coef_of = [&](int t) -> float {
        const int col = t * FB_BN + pt;
        const int part = col >= p.Bp ? 1 : 0;
        const int idx = col - part * p.Bp;
        return idx < p.B ? __ldg((part ? cr_this : cr_other) + idx) * p.inv_tau : 0.f;
      };
      const int j0 = ((kglob & 1u) == static_cast<uint32_t>(g)) ? 0 : 1;    // this pair's first valid tile of the unit
      if (j0 < vt.n && pt < FB_BN) coef[pt] = coef_of(vt.tile(j0));
      named_bar_sync(1 + g, FB_PAIR_THREADS);
      int jj = 0;
      for (int j = j0; j < vt.n; j += 2, ++jj) {
        const int t = vt.tile(j);
        // the next tile's coefficients travel from global memory while this tile is consumed
        const bool has_next = j + 2 < vt.n;
        float nxt = 0.f;
        if (has_next && pt < FB_BN) nxt = coef_of(vt.tile(j + 2));
        const int col0 = t * FB_BN + h * 32;                 // this warpgroup's 32 columns
        const int part = col0 >= p.Bp ? 1 : 0;
        const int idx0 = col0 - part * p.Bp;
        const float* cs = coef + (jj & 1) * FB_BN + h * 32;
        mbar_wait(bar(8 + g), ph);
        tc_fence_after();
        uint32_t r[32];
        SNAG_TMEM_LD32(tmem_base + lane_base + FB_TMEM_S + g * FB_BN + h * 32, r);
        SNAG_TMEM_WAIT32(r);
        tc_fence_before();
        mbar_arrive(bar(10 + g));                            // S stage free: the MMA1 two tiles ahead may overwrite it
        uint32_t w[16];
        // plain strip (warp-uniform): all 128 anchors and all 32 columns valid, no diagonal element inside
        const bool plain = (gr0 + FB_BM <= p.B) && (idx0 + 32 <= p.B) && (idx0 + 31 < gr0 || idx0 > gr0 + FB_BM - 1);
        if (plain) {
#pragma unroll
          for (int q4 = 0; q4 < 32; q4 += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cs + q4);       // broadcast read
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const float e0 = ex2_approx(__fmaf_rn(__uint_as_float(r[q4 + e]), p.scale_log2, nb));
              const float e1 = ex2_approx(__fmaf_rn(__uint_as_float(r[q4 + e + 1]), p.scale_log2, nb));
              const __nv_bfloat162 hv = __floats2bfloat162_rn((cc[e] + cr) * e0, (cc[e + 1] + cr) * e1);
              w[(q4 + e) >> 1] = *reinterpret_cast<const uint32_t*>(&hv);
            }
          }
        } else {
#pragma unroll
          for (int q4 = 0; q4 < 32; q4 += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cs + q4);
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              float v[2];
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int idx = idx0 + q4 + e + hh;
                const float E = ex2_approx(__fmaf_rn(__uint_as_float(r[q4 + e + hh]), p.scale_log2, nb));
                float gv = (cr + cc[e + hh]) * E;
                if (idx == gr) gv = part ? 0.f : gv - dg;
                if (!ok || idx >= p.B) gv = 0.f;
                v[hh] = gv;
              }
              const __nv_bfloat162 hv = __floats2bfloat162_rn(v[0], v[1]);
              w[(q4 + e) >> 1] = *reinterpret_cast<const uint32_t*>(&hv);
            }
          }
        }
        mbar_wait(bar(14 + g), ph ^ 1);                      // the MMA2 of this pair's previous tile has consumed the P buffer
        tc_fence_after();
        FB_TMEM_ST16(tmem_base + lane_base + FB_TMEM_P + g * 32 + h * 16, w);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar(12 + g));
        ph ^= 1;
        if (has_next && pt < FB_BN) coef[((jj + 1) & 1) * FB_BN + pt] = nxt;
        named_bar_sync(1 + g, FB_PAIR_THREADS);              // next coefficients visible; this tile's no longer read
      }
      kglob += static_cast<uint32_t>(vt.n);
      mbar_wait(bar(16), dzph);
      dzph ^= 1;
      tc_fence_after();
      for (int c = wg * 32; c < Dpad; c += 32 * (FB_EPI_THREADS / 128)) {    // warpgroup w takes every fourth 32-column strip
        uint32_t r[32];
        SNAG_TMEM_LD32(tmem_base + lane_base + FB_TMEM_DZ + c, r);
        SNAG_TMEM_WAIT32(r);
#pragma unroll
        for (int hh = 0; hh < 4; ++hh)
          st_global_256(orow + c + 8 * hh, make_uint4(r[8 * hh], r[8 * hh + 1], r[8 * hh + 2], r[8 * hh + 3]),
                        make_uint4(r[8 * hh + 4], r[8 * hh + 5], r[8 * hh + 6], r[8 * hh + 7]));
      }
      tc_fence_before();
      mbar_arrive(bar(17));                                  // dZ drained: the next unit may start accumulating
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cwarp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
int make_operand_map(CUtensorMap* m, const __nv_bfloat16* ptr, long long rows, int Dpad, int box_rows);   // sim_kernels.cu

// column splits per (problem, side, row block): the fewest that fill the persistent grid's waves to >= 90 %, each split
// keeping at least 8 tiles (its 128 x Dpad partial write and pipeline fill must stay small next to its MMAs)
int icl_bwd_fused_splits(int n_prob, int B, int Bp, int row_blocks) {
  if (n_prob < 1 || B < 1 || Bp < B || (Bp % 256) || row_blocks < 1) return 1;
  const int sms = num_sms();
  const long long base_units = 2ll * n_prob * row_blocks;
  const int n_tiles = 2 * Bp / FB_BN;
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= 16; ++ns) {
    const int tps = (n_tiles + ns - 1) / ns;
    if (ns > 1 && tps < 8) break;
    if ((n_tiles + tps - 1) / tps != ns) continue;           // would leave an empty split
    const long long units = base_units * ns;
    const double eff = static_cast<double>(units) / (static_cast<double>((units + sms - 1) / sms) * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ns; }
    if (eff >= 0.9) break;
  }
  return best;
}

int launch_icl_bwd_fused(int n_prob, const __nv_bfloat16* const* S3, const float* const* cr_a, const float* const* cr_b,
                         const float* const* dg, float* const* dz_a, float* const* dz_b, int B, int Bp, int rb0,
                         int row_blocks, int Dpad, float inv_tau, int nsplit, long long part_stride, cudaStream_t st) {
  if (n_prob < 1 || n_prob > FB_MAX_PROB || !S3 || !cr_a || !cr_b || !dg || !dz_a || !dz_b) return SNAG_ERR_ARG;
  if (B <= 0 || Bp < B || (Bp % 256) != 0) return SNAG_ERR_ARG;
  if (rb0 < 0 || row_blocks < 1 || (rb0 + row_blocks) * FB_BM > Bp) return SNAG_ERR_ARG;
  if (Dpad <= 0 || (Dpad % FB_BK) != 0 || Dpad > FB_MAX_KB * FB_BK) return SNAG_ERR_SHAPE;
  if (!device_is_sm100()) return SNAG_ERR_DEVICE;
  const int n_tiles = 2 * Bp / FB_BN;
  if (nsplit < 1 || nsplit > n_tiles) return SNAG_ERR_ARG;
  if (nsplit > 1 && part_stride < static_cast<long long>(row_blocks) * FB_BM * Dpad) return SNAG_ERR_ARG;
  FbParams p{};
  p.n_prob = n_prob; p.B = B; p.Bp = Bp; p.kblocks = Dpad / FB_BK;
  p.rb0 = rb0;
  p.row_blocks = row_blocks;
  p.tiles_per_split = (n_tiles + nsplit - 1) / nsplit;
  p.nsplit = (n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  if (p.nsplit != nsplit) return SNAG_ERR_SHAPE;             // the caller sized its partial buffers for nsplit
  p.n_tiles = n_tiles;
  const long long units = 2ll * n_prob * p.row_blocks * p.nsplit;
  if (units > 0x7fffffffll) return SNAG_ERR_SHAPE;
  p.n_units = static_cast<int>(units);
  p.scale_log2 = inv_tau * 1.4426950408889634f;
  p.inv_tau = inv_tau;
  p.part_stride = part_stride;
  for (int i = 0; i < n_prob; ++i) {
    if (!S3[i] || !cr_a[i] || !cr_b[i] || !dg[i] || !dz_a[i] || !dz_b[i]) return SNAG_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(cr_a[i]) | reinterpret_cast<uintptr_t>(cr_b[i])) & 15) return SNAG_ERR_ALIGN;
    if ((reinterpret_cast<uintptr_t>(dz_a[i]) | reinterpret_cast<uintptr_t>(dz_b[i]) |
         static_cast<uintptr_t>(part_stride * 4)) & 31)
      return SNAG_ERR_ALIGN;
    const int rc = make_operand_map(&p.prob[i].tm, S3[i], 3ll * Bp, Dpad, 64);
    if (rc) return rc;
    p.prob[i].cr_a = cr_a[i]; p.prob[i].cr_b = cr_b[i]; p.prob[i].dg = dg[i];
    p.prob[i].dz_a = dz_a[i]; p.prob[i].dz_b = dz_b[i];
  }
  const cudaError_t attr_err =
      cudaFuncSetAttribute(icl_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES);
  if (attr_err != cudaSuccess) return static_cast<int>(attr_err);
  const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  icl_bwd_fused_kernel<<<grid, FB_THREADS, FB_SMEM_BYTES, st>>>(p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace snag
