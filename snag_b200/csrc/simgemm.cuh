// Persistent, warp-specialised similarity contraction  S = X · Yᵀ  (bf16 in, fp32 accumulate in TMEM)
// with the consumer of S fused into the epilogue, so S never reaches HBM.
//
//   warp 0      TMA producer   : X tile [128 x 64] + Y tile [256 x 64] per k-block -> 4-stage smem ring
//   warp 1      UMMA issuer    : tcgen05.mma 128x256x16, 4 per k-block, accumulator double-buffered in TMEM
//   warp 2      TMEM allocator : 512 columns (2 accumulator stages x 256 fp32 columns)
//   warps 4..7  epilogue       : tcgen05.ld 32 lanes x 32 columns at a time, thread t <-> row t of the tile
//
// Work decomposition: a *unit* is (row block of 128 sources) x (chunk of `tiles_per_chunk` column tiles);
// per-row epilogue state (top-k list, rank counter, softmax row sum) lives in registers for the whole
// unit and is flushed once at the end. Units are ordered chunk-major so that the CTAs resident at any
// time sweep the same Y chunk (sized to stay L2 resident) while each re-reads its own X row block.
#pragma once
#include "common.cuh"

namespace snag {

constexpr int BM = 128;           // rows per tile  (UMMA M)
constexpr int BN = 256;           // columns per tile (UMMA N)
constexpr int BK = 64;            // bf16 per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;   // 512
constexpr int NUM_CTRL_THREADS = 128;
constexpr int NUM_EPI_THREADS = 128;
constexpr int NUM_THREADS = NUM_CTRL_THREADS + NUM_EPI_THREADS;
constexpr int EPI_SCRATCH_BYTES = 8192;
constexpr int BAR_BYTES = 256;
constexpr int SIM_SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + BAR_BYTES + EPI_SCRATCH_BYTES;
constexpr int KT = 16;            // per-row candidate list length kept by the top-k epilogue (k <= KT)

struct SimShape {
  int n_rows;           // valid rows of the X view
  int n_cols;           // valid rows of the Y view (= columns of S)
  int kblocks;          // Dpad / 64
  int row_blocks;       // ceil(n_rows / 128)
  int col_tiles;        // ceil(n_cols / 256)
  int tiles_per_chunk;
  int n_chunks;         // ceil(col_tiles / tiles_per_chunk)
  int n_units;          // row_blocks * n_chunks
};

struct EpiCtx {
  int et;        // epilogue thread 0..127 == row within the tile == TMEM lane
  int lane;      // lane in warp
  int row;       // row index inside the X view
  bool row_ok;   // row < n_rows
  int rb, chunk; // unit coordinates
  float* scratch;  // EPI_SCRATCH_BYTES of shared memory private to the epilogue warpgroup
};

template <class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sim_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const SimShape shp,
           const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (2 * STAGES + ACC_STAGES + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + STAGES * STAGE_BYTES +
                                                                           8 * (2 * STAGES + 2 * ACC_STAGES));
  float* scratch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = blockIdx.x; u < shp.n_units; u += gridDim.x) {
        const int rb = u % shp.row_blocks, ch = u / shp.row_blocks;
        const int ct0 = ch * shp.tiles_per_chunk;
        const int ct1 = min(ct0 + shp.tiles_per_chunk, shp.col_tiles);
        for (int ct = ct0; ct < ct1; ++ct) {
          for (int kb = 0; kb < shp.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            const uint32_t sa = base + stage * STAGE_BYTES;
            tma_load_2d(sa, &tmX, full_bar(stage), kb * BK, rb * BM);
            tma_load_2d(sa + A_STAGE_BYTES, &tmY, full_bar(stage), kb * BK, ct * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int u = blockIdx.x; u < shp.n_units; u += gridDim.x) {
        const int ch = u / shp.row_blocks;
        const int ct0 = ch * shp.tiles_per_chunk;
        const int ct1 = min(ct0 + shp.tiles_per_chunk, shp.col_tiles);
        for (int ct = ct0; ct < ct1; ++ct) {
          mbar_wait(tempty_bar(as), aphase ^ 1);   // epilogue has drained this accumulator stage
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + as * BN;
          for (int kb = 0; kb < shp.kblocks; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = base + stage * STAGE_BYTES;
            const uint64_t adesc = make_sdesc_k128(sa);
            const uint64_t bdesc = make_sdesc_k128(sa + A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // +32 bytes (encoded >>4 -> +2) per 16-element K step inside the 128-byte swizzle row
              umma_bf16_ss(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));         // smem slot free once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(tfull_bar(as));               // accumulator complete -> epilogue
          if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp >= NUM_CTRL_THREADS / 32) {
    // ------------------------------------------------------------------ epilogue warpgroup
    EpiCtx cx;
    cx.et = threadIdx.x - NUM_CTRL_THREADS;
    cx.lane = lane;
    cx.scratch = scratch;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t as = 0, aphase = 0;
    uint32_t tile_seq = 0;
    for (int u = blockIdx.x; u < shp.n_units; u += gridDim.x) {
      cx.rb = u % shp.row_blocks;
      cx.chunk = u / shp.row_blocks;
      cx.row = cx.rb * BM + cx.et;
      cx.row_ok = cx.row < shp.n_rows;
      const int ct0 = cx.chunk * shp.tiles_per_chunk;
      const int ct1 = min(ct0 + shp.tiles_per_chunk, shp.col_tiles);
      typename Epi::State st;
      Epi::unit_begin(ep, shp, cx, st);
      for (int ct = ct0; ct < ct1; ++ct, ++tile_seq) {
        const int buf = tile_seq & 1;
        Epi::tile_begin(ep, shp, cx, st, ct, buf);   // stage per-column vectors in smem (double-buffered)
        named_bar_sync(1, NUM_EPI_THREADS);
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + as * BN;
        uint32_t r0[32], r1[32];
        // software pipeline: the load of strip c+1 is in flight while strip c is consumed
        SNAG_TMEM_LD32(taddr, r0);
#pragma unroll
        for (int c = 0; c < BN / 32; c += 2) {
          SNAG_TMEM_WAIT32(r0);
          SNAG_TMEM_LD32(taddr + (c + 1) * 32, r1);
          Epi::chunk(ep, shp, cx, st, ct, c, r0, buf);
          SNAG_TMEM_WAIT32(r1);
          if (c + 2 < BN / 32) SNAG_TMEM_LD32(taddr + (c + 2) * 32, r0);
          Epi::chunk(ep, shp, cx, st, ct, c + 1, r1, buf);
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(as));
        if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        Epi::tile_end(ep, shp, cx, st, ct, buf);
      }
      Epi::unit_end(ep, shp, cx, st);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================
// Arithmetic shared by the evaluation epilogues: the reference's op order, one rounding per op.
//   d    = clamp((xn_i + yn_j) - 2 s_ij, 0)          src/utils.py:210-218
//   c    = 1 - d                                      main.py:393 (argument of csls_sim)
//   csls = (2 c - nv1_i) - nv2_j                      src/utils.py:433-434
//   dist = 1 - csls                                   main.py:393
// 2*s and 2*c are exact, so the FMAs below round exactly like the reference's separate mul and sub.
// ================================================================================================
__device__ __forceinline__ float sqdist_from_dot(float s, float xn, float yn) {
  const float t = __fadd_rn(xn, yn);
  return fmaxf(__fmaf_rn(-2.0f, s, t), 0.0f);
}
__device__ __forceinline__ float csls_dist_from_c(float c, float nv1, float nv2) {
  const float u = __fmaf_rn(2.0f, c, -nv1);
  const float v = __fsub_rn(u, nv2);
  return __fsub_rn(1.0f, v);
}

// ------------------------------------------------------------------------------------------------
// Epilogue: write S (mode 0) or the squared-L2 distance (mode 1) — drop-in pairwise_distances
// ------------------------------------------------------------------------------------------------
struct EpiWrite {
  struct Params {
    float* out;        // [n_rows, ld]
    long long ld;
    const float* xn;   // [n_rows]  (mode 1)
    const float* yn;   // [n_cols]  (mode 1)
    int mode;
  };
  struct State {
    float xn;
  };
  static __device__ __forceinline__ void unit_begin(const Params& p, const SimShape&, const EpiCtx& cx, State& st) {
    st.xn = (p.mode == 1 && cx.row_ok) ? p.xn[cx.row] : 0.f;
  }
  static __device__ __forceinline__ void tile_begin(const Params& p, const SimShape& shp, const EpiCtx& cx, State&,
                                                    int ct, int buf) {
    float* yn_s = cx.scratch + buf * BN;
    for (int j = cx.et; j < BN; j += NUM_EPI_THREADS) {
      const int col = ct * BN + j;
      yn_s[j] = (p.mode == 1 && col < shp.n_cols) ? p.yn[col] : 0.f;
    }
  }
  static __device__ __forceinline__ void chunk(const Params& p, const SimShape& shp, const EpiCtx& cx, State& st,
                                               int ct, int c, const uint32_t (&r)[32], int buf) {
    const float* yn_s = cx.scratch + buf * BN + c * 32;
    const int col0 = ct * BN + c * 32;
    if (!cx.row_ok) return;
    float* orow = p.out + static_cast<long long>(cx.row) * p.ld + col0;
    float v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const float s = __uint_as_float(r[q]);
      v[q] = (p.mode == 1) ? sqdist_from_dot(s, st.xn, yn_s[q]) : s;
    }
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) && (col0 + 32 <= shp.n_cols);
    if (vec_ok) {
#pragma unroll
      for (int q = 0; q < 32; q += 4)
        *reinterpret_cast<float4*>(orow + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (col0 + q < shp.n_cols) orow[q] = v[q];
    }
  }
  static __device__ __forceinline__ void tile_end(const Params&, const SimShape&, const EpiCtx&, State&, int, int) {}
  static __device__ __forceinline__ void unit_end(const Params&, const SimShape&, const EpiCtx&, State&) {}
};

// ------------------------------------------------------------------------------------------------
// Epilogue: per-row top-KT of c_ij = 1 - d_ij over the unit's columns (CSLS neighbourhood, sweep 1).
// Output: part[chunk][row][KT], ascending, -inf padded. A merge kernel reduces over chunks.
// ------------------------------------------------------------------------------------------------
struct EpiRowTopK {
  struct Params {
    const float* xn;   // [n_rows]
    const float* yn;   // [n_cols]
    float* part;       // [n_chunks][n_rows][KT]
  };
  struct State {
    float xn;
    float top[KT];     // ascending: top[0] is the current KT-th largest (the admission threshold)
  };
  static __device__ __forceinline__ void unit_begin(const Params& p, const SimShape&, const EpiCtx& cx, State& st) {
    st.xn = cx.row_ok ? p.xn[cx.row] : 0.f;
#pragma unroll
    for (int t = 0; t < KT; ++t) st.top[t] = -INFINITY;
  }
  static __device__ __forceinline__ void tile_begin(const Params& p, const SimShape& shp, const EpiCtx& cx, State&,
                                                    int ct, int buf) {
    float* yn_s = cx.scratch + buf * BN;
    for (int j = cx.et; j < BN; j += NUM_EPI_THREADS) {
      const int col = ct * BN + j;
      // out-of-range columns get yn = +inf  ->  d = +inf, c = -inf: never admitted
      yn_s[j] = (col < shp.n_cols) ? p.yn[col] : INFINITY;
    }
  }
  static __device__ __forceinline__ void chunk(const Params&, const SimShape&, const EpiCtx& cx, State& st, int,
                                               int c, const uint32_t (&r)[32], int buf) {
    const float* yn_s = cx.scratch + buf * BN + c * 32;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const float d = sqdist_from_dot(__uint_as_float(r[q]), st.xn, yn_s[q]);
      const float cval = __fsub_rn(1.0f, d);
      if (cval > st.top[0]) {
        st.top[0] = cval;
#pragma unroll
        for (int t = 0; t < KT - 1; ++t) {
          const float lo = fminf(st.top[t], st.top[t + 1]);
          const float hi = fmaxf(st.top[t], st.top[t + 1]);
          st.top[t] = lo;
          st.top[t + 1] = hi;
        }
      }
    }
  }
  static __device__ __forceinline__ void tile_end(const Params&, const SimShape&, const EpiCtx&, State&, int, int) {}
  static __device__ __forceinline__ void unit_end(const Params& p, const SimShape& shp, const EpiCtx& cx, State& st) {
    if (!cx.row_ok) return;
    float4* o = reinterpret_cast<float4*>(p.part + (static_cast<long long>(cx.chunk) * shp.n_rows + cx.row) * KT);
#pragma unroll
    for (int t = 0; t < KT; t += 4) o[t / 4] = make_float4(st.top[t], st.top[t + 1], st.top[t + 2], st.top[t + 3]);
  }
};

// ------------------------------------------------------------------------------------------------
// Epilogue: rank counting (sweep 2). For every pair (i, j) of the unit, with dist = CSLS distance:
//   cnt_row[i] += [dist < g_i] + [dist == g_i and gid(j) < gid(i)]      (j != i)     -> l2r rank of pair i
//   cnt_col[j] += [dist < g_j] + [dist == g_j and gid(i) < gid(j)]      (i != j)     -> r2l rank of pair j
// i.e. the position of the ground truth in a stable ascending sort (main.py:400-411, 422-429).
// Optionally tracks the 3 nearest columns per row (the ret1..ret3 of the prediction CSV, main.py:411).
// ------------------------------------------------------------------------------------------------
template <bool kTop3>
struct EpiRank {
  struct Params {
    const float* xn;     // [n_rows]
    const float* yn;     // [n_cols]
    const float* nv1;    // [n_rows]
    const float* nv2;    // [n_cols]
    const float* g_row;  // [n_rows]  dist of pair(row gid)
    const float* g_col;  // [n_cols]  dist of pair(col gid)
    int row_gid0;        // global pair id of view row 0
    int col_gid0;        // global pair id of view column 0
    int* cnt_row;        // [n_rows]  (atomically accumulated, caller zeroes)
    int* cnt_col;        // [n_cols]
    float* top3_val;     // [n_chunks][n_rows][4] (kTop3) ascending distance
    int* top3_idx;       // [n_chunks][n_rows][4] (kTop3) column gid
    int use_csls;        // 0: rank on the plain squared distance d (args.csls False, main.py:392)
  };
  struct State {
    float xn, nv1, g;
    int gid;
    int cnt;
    int colcnt[BN / 32];
    float t3v[3];
    int t3i[3];
  };
  static __device__ __forceinline__ void unit_begin(const Params& p, const SimShape&, const EpiCtx& cx, State& st) {
    st.xn = cx.row_ok ? p.xn[cx.row] : 0.f;
    st.nv1 = cx.row_ok ? p.nv1[cx.row] : 0.f;
    // invalid rows: g = -inf never counts anything on the row side; the column side masks by row_ok
    st.g = cx.row_ok ? p.g_row[cx.row] : -INFINITY;
    st.gid = p.row_gid0 + cx.row;
    st.cnt = 0;
    if (kTop3) {
#pragma unroll
      for (int t = 0; t < 3; ++t) { st.t3v[t] = INFINITY; st.t3i[t] = 0x7fffffff; }
    }
  }
  static __device__ __forceinline__ void tile_begin(const Params& p, const SimShape& shp, const EpiCtx& cx, State& st,
                                                    int ct, int buf) {
    float* s = cx.scratch + buf * (3 * BN);
    for (int j = cx.et; j < BN; j += NUM_EPI_THREADS) {
      const int col = ct * BN + j;
      const bool ok = col < shp.n_cols;
      s[j] = ok ? p.yn[col] : INFINITY;        // d = inf -> dist = +inf: never smaller than anything
      s[BN + j] = ok ? p.nv2[col] : 0.f;
      s[2 * BN + j] = ok ? p.g_col[col] : -INFINITY;
    }
#pragma unroll
    for (int q = 0; q < BN / 32; ++q) st.colcnt[q] = 0;
  }
  static __device__ __forceinline__ void chunk(const Params& p, const SimShape&, const EpiCtx& cx, State& st, int ct,
                                               int c, const uint32_t (&r)[32], int buf) {
    const float* s = cx.scratch + buf * (3 * BN) + c * 32;
    const int cgid0 = p.col_gid0 + ct * BN + c * 32;
    // relation of this 32-column strip to the rows of the tile (global pair ids), warp-uniform
    const int rgid_lo = p.row_gid0 + cx.rb * BM, rgid_hi = rgid_lo + BM - 1;
    const bool all_cols_below = (cgid0 + 31) < rgid_lo;   // every gid(j) < every gid(i)
    const bool all_cols_above = cgid0 > rgid_hi;          // every gid(j) > every gid(i)
    int cc = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const float d = sqdist_from_dot(__uint_as_float(r[q]), st.xn, s[q]);
      const float dist = p.use_csls ? csls_dist_from_c(__fsub_rn(1.0f, d), st.nv1, s[BN + q]) : d;
      const float gj = s[2 * BN + q];
      bool prow, pcol;
      if (all_cols_below) {
        prow = dist <= st.g;
        pcol = dist < gj;
      } else if (all_cols_above) {
        prow = dist < st.g;
        pcol = dist <= gj;
      } else {
        const int jg = cgid0 + q;
        prow = (dist < st.g) || (dist == st.g && jg < st.gid);
        pcol = (dist < gj) || (dist == gj && st.gid < jg);
        if (jg == st.gid) { prow = false; pcol = false; }
      }
      pcol = pcol && cx.row_ok;
      st.cnt += prow ? 1 : 0;
      const int votes = __popc(__ballot_sync(0xffffffffu, pcol));
      if (cx.lane == q) cc = votes;
      if (kTop3) {
        const int jg = cgid0 + q;
        if (dist < st.t3v[2] || (dist == st.t3v[2] && jg < st.t3i[2])) {
          st.t3v[2] = dist; st.t3i[2] = jg;
#pragma unroll
          for (int t = 2; t > 0; --t) {
            const bool sw = (st.t3v[t] < st.t3v[t - 1]) || (st.t3v[t] == st.t3v[t - 1] && st.t3i[t] < st.t3i[t - 1]);
            if (sw) {
              const float tv = st.t3v[t]; st.t3v[t] = st.t3v[t - 1]; st.t3v[t - 1] = tv;
              const int ti = st.t3i[t]; st.t3i[t] = st.t3i[t - 1]; st.t3i[t - 1] = ti;
            }
          }
        }
      }
    }
    st.colcnt[c] += cc;   // lane q holds the votes of column c*32+q (c is a compile-time constant after unrolling)
  }
  static __device__ __forceinline__ void tile_end(const Params& p, const SimShape& shp, const EpiCtx& cx, State& st,
                                                  int ct, int) {
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      const int col = ct * BN + c * 32 + cx.lane;
      if (st.colcnt[c] != 0 && col < shp.n_cols) atomicAdd(p.cnt_col + col, st.colcnt[c]);
    }
  }
  static __device__ __forceinline__ void unit_end(const Params& p, const SimShape& shp, const EpiCtx& cx, State& st) {
    if (!cx.row_ok) return;
    if (st.cnt != 0) atomicAdd(p.cnt_row + cx.row, st.cnt);
    if (kTop3) {
      const long long o = (static_cast<long long>(cx.chunk) * shp.n_rows + cx.row) * 4;
      *reinterpret_cast<float4*>(p.top3_val + o) = make_float4(st.t3v[0], st.t3v[1], st.t3v[2], 0.f);
      *reinterpret_cast<int4*>(p.top3_idx + o) = make_int4(st.t3i[0], st.t3i[1], st.t3i[2], 0);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Epilogue: in-batch contrastive row sums (ICL forward, model/SNAG_loss.py:98-126).
// X view = one side of the batch (Bp rows, B valid); Y view = [other side | same side], 2*Bp rows.
// Column c -> part p = c / Bp, idx = c - p*Bp; valid iff idx < B. p == 1 and idx == row is the
// self-similarity the reference kills with -1e9; p == 0 and idx == row is the positive logit.
// With unit-norm rows logits are bounded by 1/tau, so exp(logit - 1/tau) needs no running max.
//   rowsum_part[chunk][row] = sum_j exp2(s_ij * (log2e/tau) - log2e/tau)
//   pos[row]                = s_{row,row} of part 0
// ------------------------------------------------------------------------------------------------
struct EpiIclFwd {
  struct Params {
    float scale_log2;   // log2(e) / tau
    int B;              // valid rows per part
    int Bp;             // padded rows per part (multiple of 256)
    float* rowsum_part; // [n_chunks][Bp]
    float* pos;         // [Bp]
  };
  struct State {
    float sum;
  };
  static __device__ __forceinline__ void unit_begin(const Params&, const SimShape&, const EpiCtx&, State& st) {
    st.sum = 0.f;
  }
  static __device__ __forceinline__ void tile_begin(const Params&, const SimShape&, const EpiCtx&, State&, int, int) {}
  static __device__ __forceinline__ void chunk(const Params& p, const SimShape&, const EpiCtx& cx, State& st, int ct,
                                               int c, const uint32_t (&r)[32], int) {
    const int col0 = ct * BN + c * 32;
    const int part = col0 >= p.Bp ? 1 : 0;
    const int idx0 = col0 - part * p.Bp;
    if (idx0 >= p.B) return;                              // strip entirely in the padding (warp-uniform)
    const bool plain = (idx0 + 32 <= p.B) && (idx0 + 31 < cx.rb * BM || idx0 > cx.rb * BM + BM - 1);
    const float nb = -p.scale_log2;
    if (plain) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) acc += ex2_approx(__fmaf_rn(__uint_as_float(r[q]), p.scale_log2, nb));
      st.sum += acc;
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int idx = idx0 + q;
        const float s = __uint_as_float(r[q]);
        float e = ex2_approx(__fmaf_rn(s, p.scale_log2, nb));
        if (idx >= p.B) e = 0.f;
        if (idx == cx.row) {
          if (part == 1) e = 0.f;
          else if (cx.row_ok) p.pos[cx.row] = s;
        }
        st.sum += e;
      }
    }
  }
  static __device__ __forceinline__ void tile_end(const Params&, const SimShape&, const EpiCtx&, State&, int, int) {}
  static __device__ __forceinline__ void unit_end(const Params& p, const SimShape&, const EpiCtx& cx, State& st) {
    if (cx.row < p.Bp) p.rowsum_part[static_cast<long long>(cx.chunk) * p.Bp + cx.row] = st.sum;
  }
};

}  // namespace snag
