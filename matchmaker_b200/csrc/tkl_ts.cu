// TKL (SIGIR'20) window scores on TMA + tcgen05: the cosine of every (query row, document position) pair comes from
// the tensor cores with fp32-grade accuracy, the RBF activations, the sliding-window sums, the learned saturation
// and the dense layer are fused behind it; only [B, W] window scores are written.
//
// Reference arithmetic: matchmaker/models/published/sigir20_tkl.py:180-252 (the reference materialises
// [B, Lq, C*40, K] = 450 MB at BASELINE config 5 and re-reads it 30 times through two unfold reductions).
//
// Work decomposition.  A document is C chunk slots of 40 positions; slot c is either a packed chunk
// (slot_to_packed >= 0) or all padding.  Three slots = one TILE of 120 positions = 60 position pairs = 4 window
// BLOCKS of 15 pairs.  A window (30 positions, stride 2) is 15 consecutive pairs, so with blocks of 15 pairs every
// window is a block SUFFIX plus the next block's PREFIX: two running sums per (query row, kernel) instead of 15
// re-additions, and -- unlike prefix differences -- only additions of non-negative terms, so empty windows stay
// exactly empty (the -9900 sentinel of sigir20_tkl.py:257 keys on exact zeros).  tkl_plan_kernel counts the tiles of
// every document that can hold a non-zero window and prefix-sums them on the device; CTA x of the persistent grid
// takes the x-th equal share of that global tile sequence (documents are split where the share ends; a share that
// starts inside a document first replays the previous tile's last block -- its "halo" -- to rebuild the suffix sums).
// Windows outside every share are exactly 0 and come from one cudaMemsetAsync.
//
// Per CTA (768 threads = 6 warpgroups, registers re-dealt with setmaxnreg), same operand pipeline as
// kernel_pool_ts.cu (x = hi + lo with hi = x & 0xffffe000; [Qhi;Qlo] stacked along the UMMA N dimension, the
// document operand written to TENSOR MEMORY by the convert warps):
//   warp 0      TMA producer: per 32-column k-chunk up to three [40 x 32] chunk boxes + one [40 x 32] query box
//   warp 1      tcgen05.mma kind::tf32, M = 128 (120 used), N = 80 = [40 hi | 40 lo], A from TMEM, 3 accumulators
//   warps 2-3   query convert (hi / lo B operand, query norms)
//   warps 4-7   document convert, one thread per position (hi / lo straight into TMEM, position norms)
//   warps 8-23  epilogue.  Phase A: thread = position, 10 query columns per warp: cosine tile to shared memory (masked
//               positions -> a sentinel whose activations are exactly 0).  Phase B: thread = (query row i, kernel k),
//               walks the tile's 60 pairs in registers: activation pair sums, block prefix / suffix, window sum,
//               saturation (per-document table indexed by the window's token count), w_k * T; a 16-value transposed
//               butterfly per block reduces over the warp, 60 threads add the per-warp partials in a fixed order.
//
// The window "length" of sigir20_tkl.py:210 counts positions whose K activations do not all vanish.  When every
// cosine in [-1, 1] activates at least one kernel (checked on the device by the plan kernel: true for every kernel
// set the reference ships) that is the count of unmasked positions, which is what this kernel uses; otherwise the
// plan says so and the FFMA kernel in tkl.cu, which tests the activations themselves, runs instead.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "host_util.cuh"
#include "masks.cuh"
#include "ptx.cuh"
#include "tkl.cuh"

namespace mmb {

namespace {

constexpr int kThreads = 768;
constexpr int kChunk = 40, kWindow = 30;
constexpr int kTileSlots = 3, kTileRows = kTileSlots * kChunk;   // 120 positions
constexpr int kTilePairs = kTileRows / 2;                       // 60
constexpr int kBlk = 15, kBlocks = kTilePairs / kBlk;           // 4 blocks of 15 pairs
constexpr int kMaxLq = 40;
constexpr int kNq = 2 * kMaxLq;           // UMMA N: hi columns 0..39, lo columns 40..79
constexpr int kMaxRaw = 6;
constexpr int kOps = 4;
constexpr int kAcc = 3;
constexpr int kNormRing = kAcc + kOps + 1;
constexpr int kAccCol0 = kOps * 64;       // TMEM: [0, 256) A ring (4 x (32 hi + 32 lo)), [256, 496) 3 accumulators of 80
constexpr int kDxBytes = 128 * 128;       // [128 rows][32 fp32], rows 0..119 written
constexpr int kSlotBytes = kChunk * 128;  // one chunk's 40 rows of a k-chunk
constexpr int kQxBytes = kMaxLq * 128;
constexpr int kRawBytes = kDxBytes + kQxBytes;   // 21 KB
constexpr int kQopBytes = kNq * 128;      // B operand: rows 0-39 Q hi, rows 40-79 Q lo
constexpr int kEpiWarps = 16, kEpiThreads = kEpiWarps * 32;
constexpr int kFirstDocWarp = 4, kFirstEpiWarp = 8;
constexpr int kReleaseArrivals = 4 + 64;   // lane 0 of each document convert warp + every lane of the two query warps
// The re-deal must fit the registers the CTA was LAUNCHED with (768 threads x 80 = 61 440), not the SM's 64 K: a
// setmaxnreg.inc beyond that pool never returns.  Only the light warpgroup gives registers back; the others keep 80.
constexpr int kRegsLight = 56;   // convert and epilogue warps keep the 80 registers of the launch
constexpr int kCsStride = 122;            // floats per QUERY ROW of the cosine tile cs[i][position]: even (8-byte pair loads), 122 mod 32 = 26
                                          // puts the 3-4 query rows a warp reads at once on distinct bank pairs
constexpr int kSatStride = 33;            // table row stride (token counts 0..30)
constexpr float kSentinel = 1.0e6f;
constexpr float kTinyNorm = 1e-13f;
constexpr float kClamp = 1e-10f;

struct TsShared {
  uint64_t raw_full[kMaxRaw];
  uint64_t raw_empty[kMaxRaw];
  uint64_t op_full[kOps];
  uint64_t op_empty[kOps];
  uint64_t accfull[kAcc];
  uint64_t accempty[kAcc];
  uint32_t tmem_base;
  uint32_t pad;
  // norms travel from the convert warps to the epilogue in their own ring: the convert warps run up to kOps k-chunks
  // ahead of the MMA warp, which runs up to kAcc tiles ahead of the epilogue -- with a single k-chunk per tile
  // (D <= 32) that is kAcc + kOps tiles, so a ring indexed by the accumulator slot would be overwritten early
  float ss_d[kNormRing][128];
  float rs_q[kNormRing][kMaxLq];
  float red[kMaxLq];          // sat_emb_reduce1(q_i)
  float qm[kMaxLq];
  float sp[16];
  alignas(16) uint16_t lenw[kBlocks][16]; // 16 x token count of the window ending at each pair of the tile (byte offset into a sat row),
                              // 15 per block in a 32-byte row: two 16-byte loads per block
  float dmring[256];          // unmasked flag of the document's positions, indexed by position & 255
  float part[kEpiWarps][64];  // per-warp partial window scores
  float4 sat[kMaxLq * kSatStride];   // (sat1 * gate, sat2, sat3 * gate, -) per (query row, token count)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void split4(const float4 v, uint32_t* hi, uint32_t* lo) {
  hi[0] = __float_as_uint(v.x) & 0xffffe000u; lo[0] = __float_as_uint(v.x - __uint_as_float(hi[0]));
  hi[1] = __float_as_uint(v.y) & 0xffffe000u; lo[1] = __float_as_uint(v.y - __uint_as_float(hi[1]));
  hi[2] = __float_as_uint(v.z) & 0xffffe000u; lo[2] = __float_as_uint(v.z - __uint_as_float(hi[2]));
  hi[3] = __float_as_uint(v.w) & 0xffffe000u; lo[3] = __float_as_uint(v.w - __uint_as_float(hi[3]));
}

// ---------------------------------------------------------------------------------------------------------------
// plan: tiles per document (prefix sums), the "cover" test of the kernel set, and the share of every CTA.  One block.
//   plan[0] = cover, plan[1] = total tiles, plan[2 + b] = tiles before document b (b = 0..B),
//   plan[3 + B + b] = cost before document b (b = 0..B), plan[4 + 2B + x] = first tile of CTA x (x = 0..grid).
// A tile's cost is counted in warp-tiles of epilogue work: kTileFixedCost for everything that does not depend on the
// query (convert, MMA, phase A, barriers; from the per-role profile) plus one per epilogue warp that holds an unmasked
// query row -- short queries leave most of phase B idle, so their documents' tiles are cheaper and a CTA takes more.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTileFixedCost = 17;

// exclusive prefix sums of v[0..n) in place, v[n] = total; 1024 threads, 1024 elements per round: shuffles inside the
// warp, one shared-memory hop across warps (three barriers per round instead of twenty)
__device__ __forceinline__ void block_exclusive_scan_inplace(int32_t* v, int64_t n, int* wsum, int* carry, int32_t* total_out) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) *carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + t;
    const int x = i < n ? v[i] : 0;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int c0 = *carry;
    if (i < n) v[i] = c0 + incl - x + (warp > 0 ? wsum[warp - 1] : 0);
    __syncthreads();
    if (t == 0) *carry = c0 + wsum[31];
    __syncthreads();
  }
  if (t == 0) { *total_out = *carry; v[n] = *carry; }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) tkl_plan_kernel(const int32_t* __restrict__ slot_to_packed, const void* __restrict__ q_mask,
                                                        int mask_dtype, int64_t B, int C, int Lq, const float* __restrict__ mu,
                                                        const float* __restrict__ sigma, int K, int grid, int force_cover,
                                                        int32_t* __restrict__ plan) {
  __shared__ int sums[1024];
  __shared__ int carry;
  __shared__ int total_tiles, total_cost;
  __shared__ float klo[32], khi[32];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // tkl_ts_kernel may start its prologue now (it waits for this grid's completion before it reads the plan)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < K && t < 32) {
    const float h = 11.0f * sigma[t] / sqrtf(0.5f * 1.4426950408889634f);
    klo[t] = mu[t] - h;
    khi[t] = mu[t] + h;
  }
  const int tiles_max = (C + kTileSlots - 1) / kTileSlots;
  // up to 1024 documents the two prefix arrays live in shared memory while they are built, scanned and searched (the global
  // round trips of the scans and of the 150 binary searches were a third of this kernel's 15 us); written out once at the end
  __shared__ int32_t s_tile[1025], s_cost[1025];
  const bool in_smem = B <= 1024;
  int32_t* tile_pre = in_smem ? s_tile : plan + 2;
  int32_t* cost_pre = in_smem ? s_cost : plan + 3 + B;
  int32_t* cta_start = plan + 4 + 2 * B;
  // pass 1: tiles and cost of every document: one warp per document, four documents per warp in flight (all of their slot
  // and mask words are requested before the first ballot -- one global-memory latency per batch, not four per document)
  const int qmt = q_mask ? mask_dtype : MMB200_MASK_NONE;
  const bool small = C <= 64 && Lq <= 64;
  for (int64_t b0 = (int64_t)warp * 4; b0 < B; b0 += 128) {
    bool pk0[4], pk1[4], lv0[4], lv1[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int64_t b = b0 + a;
      const bool ok = small && b < B;
      pk0[a] = ok && lane < C && slot_to_packed[b * C + lane] >= 0;
      pk1[a] = ok && lane + 32 < C && slot_to_packed[b * C + lane + 32] >= 0;
      lv0[a] = ok && lane < Lq && mask_at(q_mask, qmt, b * Lq + lane);
      lv1[a] = ok && lane + 32 < Lq && mask_at(q_mask, qmt, b * Lq + lane + 32);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int64_t b = b0 + a;
      if (b >= B) break;   // warp-uniform
      int c_last = -1, q_hi = 0;
      if (small) {
        const unsigned m0 = __ballot_sync(0xffffffffu, pk0[a]), m1 = __ballot_sync(0xffffffffu, pk1[a]);
        c_last = m1 ? 63 - __clz(m1) : (m0 ? 31 - __clz(m0) : -1);
        const unsigned q0 = __ballot_sync(0xffffffffu, lv0[a]), q1 = __ballot_sync(0xffffffffu, lv1[a]);
        q_hi = q1 ? 64 - __clz(q1) : (q0 ? 32 - __clz(q0) : 0);
      } else {
        for (int c0 = 0; c0 < C; c0 += 32) {
          const int c = c0 + lane;
          const unsigned m = __ballot_sync(0xffffffffu, c < C && slot_to_packed[b * C + c] >= 0);
          if (m) c_last = c0 + 31 - __clz(m);
        }
        for (int i0 = 0; i0 < Lq; i0 += 32) {
          const int i = i0 + lane;
          const unsigned m = __ballot_sync(0xffffffffu, i < Lq && mask_at(q_mask, qmt, b * Lq + i));
          if (m) q_hi = i0 + 32 - __clz(m);
        }
      }
      if (lane == 0) {
        // windows overlapping a packed chunk end at the latest in slot c_last + 1
        const int tiles = c_last < 0 ? 0 : min(tiles_max, (c_last + 1) / kTileSlots + 1);
        tile_pre[b] = tiles;
        cost_pre[b] = tiles * (kTileFixedCost + ((q_hi * K + 31) >> 5));
      }
    }
  }
  __syncthreads();
  block_exclusive_scan_inplace(tile_pre, B, sums, &carry, &total_tiles);
  block_exclusive_scan_inplace(cost_pre, B, sums, &carry, &total_cost);
  // pass 3: CTA x starts at the tile where the cumulative cost reaches x / grid of the total
  for (int x = t; x <= grid; x += 1024) {
    int start = total_tiles;
    if (x < grid && total_tiles > 0) {
      const long long target = (long long)total_cost * x / grid;
      int64_t lo = 0, hi = B - 1;   // last document whose cost prefix is <= target
      while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (cost_pre[mid] <= target) lo = mid; else hi = mid - 1;
      }
      const int tiles = tile_pre[lo + 1] - tile_pre[lo];
      const int dcost = cost_pre[lo + 1] - cost_pre[lo];
      const int per_tile = tiles > 0 ? dcost / tiles : 1;
      start = tile_pre[lo] + (tiles > 0 ? min(tiles, (int)((target - cost_pre[lo] + per_tile - 1) / per_tile)) : 0);
    }
    cta_start[x] = start;
  }
  if (in_smem)
    for (int64_t i = t; i <= B; i += 1024) { plan[2 + i] = s_tile[i]; plan[3 + B + i] = s_cost[i]; }
  // Cover: activation k is non-zero (ex2.approx.ftz) for |c - mu_k| * a_k <= sqrt(126); 11.0 leaves a margin.  The union of
  // the intervals [klo, khi] covers [-1.01, 1.01] iff the left end and every right end inside the range lie inside an
  // interval that extends beyond them -- one thread per end point instead of a serial sweep.
  __shared__ int uncovered;
  if (t == 0) uncovered = 0;
  __syncthreads();
  if (t <= K && t <= 32) {
    const float x = t == 0 ? -1.01f : khi[t - 1];
    if (x >= -1.01f && x < 1.01f) {
      bool ok = false;
      for (int k = 0; k < K; ++k) ok = ok || (klo[k] <= x && khi[k] > x);
      if (!ok) atomicExch(&uncovered, 1);
    }
  }
  __syncthreads();
  if (t == 0) {
    plan[1] = total_tiles;
    plan[0] = (uncovered == 0 || force_cover == 1) && force_cover != -1 ? 1 : 0;
  }
}

// The tiles a CTA walks, identically in every role: [g_begin, g_end) of the global tile sequence, preceded by a halo
// tile when the share starts inside a document.
struct TileWalk {
  const int32_t* pre;   // plan + 2
  int g, g_end;
  int b;
  int t;                // tile inside document b
  int b_end_tile;       // tiles of document b
  bool halo;
  __device__ __forceinline__ bool init(const int32_t* plan, int B, int cta, int ncta) {
    pre = plan + 2;
    const int32_t* cta_start = plan + 4 + 2 * B;   // shares of equal COST (tkl_plan_kernel), monotone in the CTA index
    g = cta_start[cta];
    g_end = cta_start[cta + 1];
    if (g >= g_end) return false;
    int lo = 0, hi = B - 1;   // last document with pre[b] <= g ...
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pre[mid] <= g) lo = mid; else hi = mid - 1;
    }
    b = lo;
    while (pre[b + 1] <= g) ++b;  // ... that has tiles (documents without tiles repeat the prefix value)
    t = g - pre[b];
    b_end_tile = pre[b + 1] - pre[b];
    halo = t > 0;
    if (halo) --t;
    return true;
  }
  __device__ __forceinline__ bool valid() const { return halo || g < g_end; }
  __device__ __forceinline__ void next() {
    if (halo) { halo = false; ++t; return; }
    ++g; ++t;
    if (g < g_end && t == b_end_tile) {
      ++b;
      while (pre[b + 1] == pre[b]) ++b;
      t = 0;
      b_end_tile = pre[b + 1] - pre[b];
    }
  }
};

// MMB200_ENABLE_PROF builds (python -m matchmaker_b200.build --prof) + MMB200_TKL_TS_PROF=1: one thread per role of CTA 0
// accumulates the cycles it spends in each wait / phase; printed by the launcher.  Compiled out of the product build.
#ifdef MMB200_ENABLE_PROF
#define TKL_T(slot, stmt)                    \
  do {                                       \
    const long long t0_ = clock64();         \
    stmt;                                    \
    pc[slot] += clock64() - t0_;             \
  } while (0)
#define TKL_MARK(slot)                       \
  do {                                       \
    const long long now_ = clock64();        \
    pc[slot] += now_ - t_mark;               \
    t_mark = now_;                           \
  } while (0)
#else
#define TKL_T(slot, stmt) stmt
#define TKL_MARK(slot) \
  do {                 \
  } while (0)
#endif

template <int SAT>
__global__ void __launch_bounds__(kThreads, 1)
tkl_ts_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c, TklParams P,
              int n_raw, int fallback_available, long long* prof) {
#ifdef MMB200_ENABLE_PROF
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long t_mark = clock64();
  const long long t_start = t_mark;
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* qring = smem;                                                    // [kOps][Qhi;Qlo]
  uint8_t* raws = smem + kOps * kQopBytes;                                  // [n_raw][Dx | Qx]
  float* cs = reinterpret_cast<float*>(raws + (size_t)n_raw * kRawBytes);   // [40 query rows][kCsStride] cosine tile
  TsShared* S = reinterpret_cast<TsShared*>(cs + kMaxLq * kCsStride);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = (P.D + 31) / 32;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_c);
    for (int s = 0; s < n_raw; ++s) { mbar_init(&S->raw_full[s], 1); mbar_init(&S->raw_empty[s], kReleaseArrivals); }
    for (int s = 0; s < kOps; ++s) { mbar_init(&S->op_full[s], kReleaseArrivals); mbar_init(&S->op_empty[s], 1); }
    for (int s = 0; s < kAcc; ++s) { mbar_init(&S->accfull[s], 1); mbar_init(&S->accempty[s], kEpiWarps); }
    fence_barrier_init();
  }
  if (threadIdx.x < 16) S->sp[threadIdx.x] = (SAT == 0 && threadIdx.x < 13) ? P.sat_params[threadIdx.x] : 0.f;
  if (warp == 1) tmem_alloc(&S->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = S->tmem_base;

  // Programmatic dependent launch: this grid is launched while the plan kernel still runs (it triggers its dependents at
  // its first instruction), so the launch latency and the prologue above overlap with it; everything below reads the plan.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const bool covered = P.plan[0] == 1;
  if (!covered && !fallback_available && blockIdx.x == 0 && threadIdx.x == 0) {
    printf("mmb200 tkl: kernel set does not cover the cosine range and the FFMA kernel cannot run this shape\n");
    __trap();
  }

  // every role walks the same tile sequence with its own copy of the iterator (set up inside the role branch, after
  // setmaxnreg, so that it lives in that role's registers); without cover nobody has work and the FFMA kernel takes over
#define TKL_WALK() TileWalk tw; const bool have_work = covered && tw.init(P.plan, (int)P.B, (int)blockIdx.x, (int)gridDim.x)

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    setmaxnreg_dec<kRegsLight>();
    TKL_WALK();
    if (lane == 0 && have_work) {
      int stage = 0;
      uint32_t phase = 0;
      for (; tw.valid(); tw.next()) {
        int pk[kTileSlots];
        int n_present = 0;
#pragma unroll
        for (int s = 0; s < kTileSlots; ++s) {
          const int c = tw.t * kTileSlots + s;
          pk[s] = (c < P.C && !(tw.halo && s < kTileSlots - 1)) ? P.slot_to_packed[(int64_t)tw.b * P.C + c] : -1;
          n_present += pk[s] >= 0 ? 1 : 0;
        }
        const uint32_t bytes = (uint32_t)(n_present * kSlotBytes + kQxBytes);
        for (int ck = 0; ck < nch; ++ck) {
          TKL_T(0, mbar_wait<true>(&S->raw_empty[stage], phase ^ 1u));
          uint8_t* st = raws + (size_t)stage * kRawBytes;
          mbar_arrive_expect_tx(&S->raw_full[stage], bytes);
#pragma unroll
          for (int s = 0; s < kTileSlots; ++s)
            if (pk[s] >= 0) tma_load_3d(&tmap_c, st + s * kSlotBytes, &S->raw_full[stage], ck * 32, 0, pk[s], kEvictFirst);
          tma_load_3d(&tmap_q, st + kDxBytes, &S->raw_full[stage], ck * 32, 0, (int)tw.b, kEvictLast);
          if (++stage == n_raw) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    setmaxnreg_dec<kRegsLight>();
    TKL_WALK();
    if (have_work) {
      const uint32_t idesc = make_idesc(kFmtTF32, 128, kNq);
      int stage = 0, acc = 0;
      uint32_t phase = 0, accphase = 0;
      for (; tw.valid(); tw.next()) {
        TKL_T(0, mbar_wait<true>(&S->accempty[acc], accphase ^ 1u));
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + (uint32_t)(kAccCol0 + acc * kNq);
        for (int ck = 0; ck < nch; ++ck) {
          const int ksteps = (min(32, P.D - ck * 32) + 7) >> 3;
          TKL_T(1, mbar_wait<true>(&S->op_full[stage], phase));
          tc_fence_after_sync();
          const uint32_t abase = tmem_base + (uint32_t)(stage * 64);
          const uint64_t b0 = make_sw128_kmajor_desc(smem_u32(qring + (size_t)stage * kQopBytes));
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < ksteps) {
                const uint64_t bq = b0 + (uint64_t)(k * 2);
                umma_tf32_ts(tmem_d, abase + (uint32_t)(k * 8), bq, idesc, (uint32_t)((ck | k) != 0));
                umma_tf32_ts(tmem_d, abase + (uint32_t)(32 + k * 8), bq, idesc, 1u);
              }
            }
            umma_commit(&S->op_empty[stage]);
            if (ck == nch - 1) umma_commit(&S->accfull[acc]);
          }
          __syncwarp();
          if (++stage == kOps) { stage = 0; phase ^= 1u; }
        }
        if (++acc == kAcc) { acc = 0; accphase ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // ------------------------------- query convert ------------------------------
    // 40 rows x 8 float4 per k-chunk = 320 float4 over 64 threads: thread qt owns 16-byte column c = qt & 7 of rows
    // (qt >> 3) + 8 j, j = 0..4
    setmaxnreg_dec<kRegsLight>();
    TKL_WALK();
    if (have_work) {
      const int qt = (warp - 2) * 32 + lane;
      const int c = qt & 7, r0 = qt >> 3;
      int rs_ = 0, os_ = 0, nr = 0;
      uint32_t rphase = 0, ophase = 0;
      for (; tw.valid(); tw.next()) {
        float ss[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int ck = 0; ck < nch; ++ck) {
          TKL_T(0, mbar_wait<true>(&S->raw_full[rs_], rphase));
          const uint8_t* xq = raws + (size_t)rs_ * kRawBytes + kDxBytes;
          float4 x[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const int row = r0 + 8 * j;
            x[j] = *reinterpret_cast<const float4*>(xq + row * 128 + ((c ^ (row & 7)) << 4));
            ss[j] = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, fmaf(x[j].z, x[j].z, fmaf(x[j].w, x[j].w, ss[j]))));
          }
          TKL_T(1, mbar_wait<true>(&S->op_empty[os_], ophase ^ 1u));
          uint8_t* qo = qring + (size_t)os_ * kQopBytes;
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const int row = r0 + 8 * j;
            uint32_t hi[4], lo[4];
            split4(x[j], hi, lo);
            const int off = row * 128 + ((c ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(qo + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(qo + kMaxLq * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          if (ck == nch - 1) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              float v = ss[j];
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              if (c == 0) S->rs_q[nr][r0 + 8 * j] = 1.0f / (sqrtf(v) + kTinyNorm);
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&S->raw_empty[rs_]);
          mbar_arrive(&S->op_full[os_]);
          if (++rs_ == n_raw) { rs_ = 0; rphase ^= 1u; }
          if (++os_ == kOps) { os_ = 0; ophase ^= 1u; }
        }
        if (++nr == kNormRing) nr = 0;
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // ------------------------------- document convert ---------------------------
    // one thread per position (TMEM lane): this kernel is bound by the SM's issue slots (ncu: 65 % issue-active, the
    // MUFU-heavy epilogue next door), not by the latency of the convert chain, so the per-chunk bookkeeping (barrier
    // waits, address arithmetic, arrivals) is paid by 4 warps instead of 8
    TKL_WALK();
    if (have_work) {
      const int qd = warp & 3;
      const int row = qd * 32 + lane;
      const int sw = row & 7;
      const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
      int rs_ = 0, os_ = 0, nr = 0;
      uint32_t rphase = 0, ophase = 0;
      for (; tw.valid(); tw.next()) {
        float4 ss4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ck = 0; ck < nch; ++ck) {
          const bool second = P.D - ck * 32 > 16;   // columns 16..31 of this chunk hold data (warp-uniform)
          TKL_T(0, mbar_wait<true>(&S->raw_full[rs_], rphase));
          const uint8_t* xrow = raws + (size_t)rs_ * kRawBytes + row * 128;
          float4 x[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const float4*>(xrow + ((c ^ sw) << 4));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 v = x[c];
            ss4.x = fmaf(v.x, v.x, ss4.x); ss4.y = fmaf(v.y, v.y, ss4.y); ss4.z = fmaf(v.z, v.z, ss4.z); ss4.w = fmaf(v.w, v.w, ss4.w);
          }
          TKL_T(1, mbar_wait<true>(&S->op_empty[os_], ophase ^ 1u));
          tc_fence_after_sync();
          const uint32_t taddr = trow + (uint32_t)(os_ * 64);
          {   // columns 0..15 of the chunk
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) split4(x[c], hi + 4 * c, lo + 4 * c);
            tmem_st_32x32b_x16(taddr, hi);
            tmem_st_32x32b_x16(taddr + 32, lo);
          }
          if (second) {   // columns 16..31: second pass over the same registers
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const float4*>(xrow + (((4 + c) ^ sw) << 4));
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 v = x[c];
              ss4.x = fmaf(v.x, v.x, ss4.x); ss4.y = fmaf(v.y, v.y, ss4.y); ss4.z = fmaf(v.z, v.z, ss4.z); ss4.w = fmaf(v.w, v.w, ss4.w);
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) split4(x[c], hi + 4 * c, lo + 4 * c);
            tmem_st_32x32b_x16(taddr + 16, hi);
            tmem_st_32x32b_x16(taddr + 48, lo);
          }
          tmem_st_wait();
          if (ck == nch - 1) S->ss_d[nr][row] = (ss4.x + ss4.y) + (ss4.z + ss4.w);
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&S->raw_empty[rs_]);
            mbar_arrive(&S->op_full[os_]);
          }
          if (++rs_ == n_raw) { rs_ = 0; rphase ^= 1u; }
          if (++os_ == kOps) { os_ = 0; ophase ^= 1u; }
        }
        if (++nr == kNormRing) nr = 0;
      }
    }
  } else {
    // ------------------------------- epilogue ------------------------------------
    TKL_WALK();
    if (have_work) {
      const int ew = warp - kFirstEpiWarp;       // 0..15
      const int et = ew * 32 + lane;             // 0..511
      const int qd = warp & 3;                   // TMEM lane quarter
      const int cg = ew >> 2;                    // query columns 10 cg .. 10 cg + 9 in phase A
      const int row = qd * 32 + lane;            // position inside the tile
      const int dmt = P.chunk_mask ? P.mask_dtype : MMB200_MASK_NONE;
      const int qmt = P.q_mask ? P.mask_dtype : MMB200_MASK_NONE;
      const int n_ik = P.Lq * P.K;
      const bool ik_live = et < n_ik;
      const int qi = ik_live ? et / P.K : 0;     // query row of this thread in phase B
      const int kk = ik_live ? et - qi * P.K : 0;
      const float a_k = sqrtf(0.5f * 1.4426950408889634f) / P.sigma[kk];
      const float nma_k = -P.mu[kk] * a_k;        // x = (c - mu) a = c a + nma
      const float w_k = ik_live ? P.dense_w[kk] : 0.f;
      const float km_k = SAT == 1 ? P.sat_params[kk] : 1.f;
      float suf[kBlk];
#pragma unroll
      for (int r = 0; r < kBlk; ++r) suf[r] = 0.f;
      int acc_slot = 0, nr = 0;
      uint32_t accphase = 0;
      // This position's mask word needs two dependent global loads (slot map, then mask).  They are issued one and two
      // tiles ahead -- the packed index of tile n + 2 and, with the index fetched a tile earlier, the mask word of tile
      // n + 1 -- at the end of phase A, so each has a whole phase B to arrive.
      auto fetch_slot = [&](int bb, int tt) -> int {
        if (row >= kTileRows) return -1;
        const int c = tt * kTileSlots + row / kChunk;
        return c < P.C ? P.slot_to_packed[(int64_t)bb * P.C + c] : -1;
      };
      auto fetch_word = [&](int pk) -> uint64_t {
        if (pk < 0) return 0;
        return dmt != MMB200_MASK_NONE ? mask_raw(P.chunk_mask, dmt, (int64_t)pk * kChunk + (row % kChunk)) : 1;
      };
      int have_ahead = 0;          // how many of the following tiles have their prefetches in flight (0, 1 or 2)
      int pk_n1 = -1, pk_n2 = -1;  // packed chunk index of this position in tiles n + 1, n + 2
      uint64_t draw_n1 = 0;
      int cur_doc = -1;
      int n_doc_warps = 0;   // epilogue warps holding an unmasked query row of the current document
      float qm_i = 0.f;

      for (; tw.valid(); tw.next()) {
        const int b = tw.b;
        const int t = tw.t;
        if (b != cur_doc) {
          // ---- new document: query mask, sat_emb_reduce1(q_i), saturation table indexed by (query row, token count)
          cur_doc = b;
          named_bar_sync(1, kEpiThreads);   // nobody still reads the previous document's table / qm
          if (et < kMaxLq) S->qm[et] = (et < P.Lq && mask_at(P.q_mask, qmt, (int64_t)b * P.Lq + et)) ? 1.f : 0.f;
          if (SAT == 0) {
            // the warp's rows ew, ew + 16, ew + 32 side by side: their loads are independent and in flight together (one
            // global-memory latency per document switch instead of one per row and 128-byte step)
            float rd[3] = {0.f, 0.f, 0.f};
            const float4* wr = reinterpret_cast<const float4*>(P.sat_red_w);
#pragma unroll 2
            for (int c4 = lane; c4 < (P.D >> 2); c4 += 32) {
              const float4 w = __ldg(wr + c4);
              float4 v[3];
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                const int i = ew + kEpiWarps * a;
                v[a] = i < P.Lq ? __ldg(reinterpret_cast<const float4*>(P.q + ((int64_t)b * P.Lq + i) * P.D) + c4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int a = 0; a < 3; ++a)
                rd[a] = fmaf(v[a].x, w.x, fmaf(v[a].y, w.y, fmaf(v[a].z, w.z, fmaf(v[a].w, w.w, rd[a]))));
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) rd[a] += __shfl_xor_sync(0xffffffffu, rd[a], o);
              if (lane == 0 && ew + kEpiWarps * a < kMaxLq) S->red[ew + kEpiWarps * a] = rd[a];
            }
          }
          named_bar_sync(1, kEpiThreads);
          if (SAT == 0) {
            const float* sp = S->sp;
            for (int e = et; e < kMaxLq * 31; e += kEpiThreads) {
              const int i = e / 31, len = e - i * 31;
              // LayerNorm over the pair (reduce(q_i), len), then three Linear(2,1) (sigir20_tkl.py:224-234); the gate
              // q_mask[i] * (len > 0) of :248 is folded into sat1 / sat3
              const float a0 = S->red[i], a1 = (float)len;
              const float mean = (a0 + a1) * 0.5f;
              const float d0 = a0 - mean, d1 = a1 - mean;
              const float rstd = rsqrtf((d0 * d0 + d1 * d1) * 0.5f + 1e-5f);
              const float y0 = d0 * rstd * sp[0] + sp[2], y1 = d1 * rstd * sp[1] + sp[3];
              const float gate = (S->qm[i] != 0.f && len > 0) ? 1.f : 0.f;
              S->sat[i * kSatStride + len] = make_float4((y0 * sp[4] + y1 * sp[5] + sp[6]) * gate,
                                                         1.0f / (y0 * sp[7] + y1 * sp[8] + sp[9]),
                                                         (y0 * sp[10] + y1 * sp[11] + sp[12]) * gate, 0.f);
            }
          }
          qm_i = S->qm[qi];   // written before the barrier above
          // query rows past the last unmasked one contribute exactly 0 to every window (the gate q_mask[i] of :248):
          // the warps that hold only such rows sit phase B out -- MSMARCO queries fill a fraction of the Lq slots
          int q_hi = 0;
          for (int i = kMaxLq - 1; i >= 0; --i)
            if (S->qm[i] != 0.f) { q_hi = i + 1; break; }
          n_doc_warps = (q_hi * P.K + 31) >> 5;
        }

        // ---- phase A: accumulator -> cosine tile -------------------------------------------------------------
        // this position's mask word was fetched while the previous tile was in phase B (two dependent global loads --
        // slot map, then mask -- that the accumulator wait no longer hides: the producers run ahead of the epilogue)
        int pk_cur;
        uint64_t draw;
        if (have_ahead == 0) {   // first tile of the share: nothing was prefetched
          pk_cur = fetch_slot(b, t);
          draw = fetch_word(pk_cur);
        } else {
          pk_cur = pk_n1;
          draw = draw_n1;
        }
        const bool present = pk_cur >= 0;
        TKL_MARK(7);   // bookkeeping between tiles, document switch
        mbar_wait<true>(&S->accfull[acc_slot], accphase);
        TKL_MARK(0);   // wait for the accumulator
        tc_fence_after_sync();
        {
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(kAccCol0 + acc_slot * kNq + 10 * cg);
          uint32_t h8[8], h2[2], l8[8], l2[2];
          tmem_ld_32x32b_x8(taddr, h8);
          tmem_ld_32x32b_x2(taddr + 8, h2);
          tmem_ld_32x32b_x8(taddr + kMaxLq, l8);
          tmem_ld_32x32b_x2(taddr + kMaxLq + 8, l2);
          tmem_ld_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->accempty[acc_slot]);
          const bool valid = present && mask_test(draw, dmt);
          if (row < kTileRows) {
            const float rsd = 1.0f / (sqrtf(S->ss_d[nr][row]) + kTinyNorm);
            const float* rq = S->rs_q[nr] + 10 * cg;
            float v[10];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = valid ? (__uint_as_float(h8[j]) + __uint_as_float(l8[j])) * rsd * rq[j] : kSentinel;
#pragma unroll
            for (int j = 0; j < 2; ++j) v[8 + j] = valid ? (__uint_as_float(h2[j]) + __uint_as_float(l2[j])) * rsd * rq[8 + j] : kSentinel;
            float* dst = cs + (10 * cg) * kCsStride + row;   // transposed: consecutive positions of one query row are adjacent
#pragma unroll
            for (int j = 0; j < 10; ++j) dst[j * kCsStride] = v[j];
            if (cg == 0) S->dmring[(t * kTileRows + row) & 255] = valid ? 1.f : 0.f;
          }
        }
        if (++acc_slot == kAcc) { acc_slot = 0; accphase ^= 1u; }
        if (++nr == kNormRing) nr = 0;
        {   // prefetches for the next two tiles (see fetch_slot above)
          TileWalk tn = tw;
          tn.next();
          if (!tn.valid()) {
            have_ahead = 0;
          } else {
            pk_n1 = have_ahead == 2 ? pk_n2 : fetch_slot(tn.b, tn.t);
            draw_n1 = fetch_word(pk_n1);   // waits for pk_n1 only on the first tile of a share
            tn.next();
            if (tn.valid()) { pk_n2 = fetch_slot(tn.b, tn.t); have_ahead = 2; } else { have_ahead = 1; }
          }
        }
        TKL_MARK(1);   // phase A
        named_bar_sync(2, kEpiThreads);
        TKL_MARK(2);   // barrier 2
        // ---- token count of the window ending at each pair of this tile (sigir20_tkl.py:210 under "cover") ----
        if (et < 8 * kTilePairs) {   // eight threads per window, four positions each, three shuffle steps
          const int wl = et >> 3, sub = et & 7;
          const int last = t * kTileRows + 2 * wl + 1;   // last position of the window ending at pair wl
          float n = 0.f;
#pragma unroll
          for (int u = sub; u < kWindow; u += 8) {
            const int pos = last - u;
            if (pos >= 0) n += S->dmring[pos & 255];
          }
          n += __shfl_xor_sync(0xffffffffu, n, 1);
          n += __shfl_xor_sync(0xffffffffu, n, 2);
          n += __shfl_xor_sync(0xffffffffu, n, 4);
          if (sub == 0) S->lenw[wl / kBlk][wl % kBlk] = (uint16_t)(16 * (int)n);
        }
        TKL_MARK(3);   // token counts
        named_bar_sync(3, kEpiThreads);
        TKL_MARK(2);
        // ---- phase B: activations, block prefix / suffix, windows ------------------------------------------------
        if (ew < n_doc_warps) {
          if (tw.halo) {
            // halo tile: only the suffix sums of its last block are wanted
            const float2* cblk = reinterpret_cast<const float2*>(cs + qi * kCsStride + 2 * kBlk * (kBlocks - 1));
#pragma unroll
            for (int r = 0; r < kBlk; ++r) {
              const float2 c = cblk[r];
              const float x0 = fmaf(c.x, a_k, nma_k), x1 = fmaf(c.y, a_k, nma_k);
              suf[r] = ex2f(-x0 * x0) + ex2f(-x1 * x1);
            }
#pragma unroll
            for (int r = kBlk - 2; r >= 0; --r) suf[r] += suf[r + 1];
          } else {
            // Windows that fall outside [0, W) (first block of a document, tail of the last tile) are computed like
            // the others -- from stale suffix sums or padding -- and dropped by the range check of the final write.
            const uint8_t* sat_row = reinterpret_cast<const uint8_t*>(S->sat + qi * kSatStride);
            for (int j = 0; j < kBlocks; ++j) {
              float tv[16];
              float pre = 0.f;
              const float2* cblk = reinterpret_cast<const float2*>(cs + qi * kCsStride + 2 * kBlk * j);
              uint32_t lw[8];   // the block's 15 window token counts: two 16-byte loads instead of 15 scalar ones
              {
                const uint4 la = *reinterpret_cast<const uint4*>(&S->lenw[j][0]), lb = *reinterpret_cast<const uint4*>(&S->lenw[j][8]);
                lw[0] = la.x; lw[1] = la.y; lw[2] = la.z; lw[3] = la.w; lw[4] = lb.x; lw[5] = lb.y; lw[6] = lb.z; lw[7] = lb.w;
              }
#pragma unroll
              for (int r = 0; r < kBlk; ++r) {
                const float2 c = cblk[r];
                const float x0 = fmaf(c.x, a_k, nma_k), x1 = fmaf(c.y, a_k, nma_k);
                const float u = ex2f(-x0 * x0) + ex2f(-x1 * x1);
                pre = r == 0 ? u : pre + u;
                const float Ssum = r < kBlk - 1 ? suf[r + 1] + pre : pre;
                suf[r] = u;   // suf[r] of the previous block was consumed by window r - 1
                const int lenb = (int)((lw[r >> 1] >> (16 * (r & 1))) & 0xffffu);
                if (SAT == 0) {
                  const float4 st = *reinterpret_cast<const float4*>(sat_row + lenb);
                  const float pw = ex2f(st.y * lg2f(fmaxf(Ssum, kClamp)));
                  tv[r] = w_k * fmaf(st.x, pw, -st.z);
                } else {
                  const float lg = 0.6931471805599453f * lg2f(fmaxf(Ssum * km_k, kClamp));
                  tv[r] = (lenb > 0 && qm_i != 0.f) ? w_k * lg : 0.f;
                }
              }
#pragma unroll
              for (int r = kBlk - 2; r >= 0; --r) suf[r] += suf[r + 1];
              tv[15] = 0.f;
              // transposed butterfly: 16 values x 32 lanes -> lane l holds the warp total of value l >> 1
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const bool up = (lane & 16) != 0;
                const float send = up ? tv[jj] : tv[jj + 8];
                const float keep = up ? tv[jj + 8] : tv[jj];
                tv[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
              }
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const bool up = (lane & 8) != 0;
                const float send = up ? tv[jj] : tv[jj + 4];
                const float keep = up ? tv[jj + 4] : tv[jj];
                tv[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const bool up = (lane & 4) != 0;
                const float send = up ? tv[jj] : tv[jj + 2];
                const float keep = up ? tv[jj + 2] : tv[jj];
                tv[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              {
                const bool up = (lane & 2) != 0;
                const float send = up ? tv[0] : tv[1];
                const float keep = up ? tv[1] : tv[0];
                tv[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              tv[0] += __shfl_xor_sync(0xffffffffu, tv[0], 1);
              const int vi = lane >> 1;   // = 8 b4 + 4 b3 + 2 b2 + b1
              if (!(lane & 1) && vi < kBlk) S->part[ew][kBlk * j + vi] = tv[0];
            }
          }
        }
        TKL_MARK(4);   // phase B
        named_bar_sync(4, kEpiThreads);
        TKL_MARK(5);   // barrier 4
        if (!tw.halo && et < kTilePairs) {
          const int w = t * kTilePairs - (kBlk - 1) + et;
          if (w >= 0 && w < P.W) {
            float pv[kEpiWarps];
#pragma unroll
            for (int ww = 0; ww < kEpiWarps; ++ww) pv[ww] = ww < n_doc_warps ? S->part[ww][et] : 0.f;   // loads in flight together
            float s = 0.f;
#pragma unroll
            for (int ww = 0; ww < kEpiWarps; ++ww) s += pv[ww];   // fixed order; the zeros of idle warps change nothing
            P.window_score[(int64_t)b * P.W + w] = s;
          }
        }
      }
    }
  }

#ifdef MMB200_ENABLE_PROF
  if (prof && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == kFirstDocWarp || warp == kFirstEpiWarp)) {
    const int role = warp == 0 ? 0 : warp == 1 ? 1 : warp == 2 ? 2 : warp == kFirstDocWarp ? 3 : 4;
    for (int i = 0; i < 8; ++i) prof[role * 9 + i] = pc[i];
    prof[role * 9 + 8] = clock64() - t_start;
  }
  if (prof && lane == 0 && (warp == kFirstEpiWarp || warp == 1)) {   // per CTA: epilogue / MMA role totals, phase B, accumulator wait
    long long* pq = prof + 45 + (long long)blockIdx.x * 4;
    if (warp == 1) pq[3] = clock64() - t_start;
    else { pq[0] = clock64() - t_start; pq[1] = pc[4]; pq[2] = pc[0]; }
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

#undef TKL_WALK

}  // namespace

int tkl_window_ts_launch(TklParams& P, const DeviceInfo& dev, cudaStream_t stream, bool* handled, int32_t** plan_out) {
  *handled = false;
  *plan_out = nullptr;
  if (P.Lq > kMaxLq || P.K > 16 || P.Lq * P.K > kEpiThreads || P.D % 4 != 0) return MMB200_OK;
  if (P.B * (int64_t)P.C >= (1ll << 31) || P.B >= (1ll << 31) - 8) return MMB200_OK;
  const size_t fixed = (size_t)kOps * kQopBytes + (size_t)kMaxLq * kCsStride * sizeof(float) + sizeof(TsShared) + 1024;
  const int n_raw = std::min<int>(kMaxRaw, (int)(((size_t)dev.max_smem_optin - fixed) / kRawBytes));
  if (n_raw < 2) return MMB200_OK;
  const size_t smem = fixed + (size_t)n_raw * kRawBytes;
  CUtensorMap tq, tc;
  {
    const uint64_t dims[3] = {(uint64_t)P.D, (uint64_t)P.Lq, (uint64_t)P.B};
    const uint64_t strides[2] = {(uint64_t)P.D * 4, (uint64_t)P.Lq * P.D * 4};
    const uint32_t box[3] = {32, (uint32_t)kMaxLq, 1};
    if (int rc = encode_tensor_map(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.q, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
      return rc;
  }
  {
    // the packed chunk count is not part of the C ABI: every index the kernel uses comes from slot_to_packed, so the
    // outer extent only has to be an upper bound (B * C slots)
    const uint64_t dims[3] = {(uint64_t)P.D, (uint64_t)kChunk, (uint64_t)(P.n_chunks > 0 ? P.n_chunks : P.B * P.C)};
    const uint64_t strides[2] = {(uint64_t)P.D * 4, (uint64_t)kChunk * P.D * 4};
    const uint32_t box[3] = {32, (uint32_t)kChunk, 1};
    if (int rc = encode_tensor_map(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.chunks, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  int32_t* plan = nullptr;
  const int grid = dev.sm_count;
  MMB_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&plan), (size_t)(2 * P.B + 6 + grid) * sizeof(int32_t), stream));
  int force = 0;
#ifdef MMB200_ENABLE_PROF
  if (const char* e = getenv("MMB200_TKL_COVER")) force = atoi(e);  // 1 / -1: force the answer of the cover test
#endif
  // the zero fill first: the window-score kernel is a programmatic dependent of the plan kernel (nothing may sit between them)
  MMB_CHECK_CUDA(cudaMemsetAsync(P.window_score, 0, (size_t)P.B * P.W * sizeof(float), stream));
  tkl_plan_kernel<<<1, 1024, 0, stream>>>(P.slot_to_packed, P.q_mask, P.mask_dtype, P.B, P.C, P.Lq, P.mu, P.sigma, P.K, grid,
                                          force, plan);
  MMB_CHECK_CUDA(cudaGetLastError());
  P.plan = plan;
  *plan_out = plan;
  const int fallback = P.segs > 0 ? 1 : 0;
  long long* prof = nullptr;
#ifdef MMB200_ENABLE_PROF
  const bool do_prof = getenv("MMB200_TKL_TS_PROF") != nullptr;
  if (do_prof) {
    MMB_CHECK_CUDA(cudaMalloc(&prof, (45 + 4 * 160) * sizeof(long long)));
    MMB_CHECK_CUDA(cudaMemset(prof, 0, (45 + 4 * 160) * sizeof(long long)));
  }
#endif
  auto launch_pdl = [&](auto kernel) -> cudaError_t {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, tq, tc, P, n_raw, fallback, prof);
  };
  static bool attr_set[2][64] = {};
  const int di = dev.device & 63;
  if (P.saturation == 0) {
    if (!attr_set[0][di]) {
      MMB_CHECK_CUDA(cudaFuncSetAttribute(tkl_ts_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dev.max_smem_optin));
      attr_set[0][di] = true;
    }
    MMB_CHECK_CUDA(launch_pdl(tkl_ts_kernel<0>));
  } else {
    if (!attr_set[1][di]) {
      MMB_CHECK_CUDA(cudaFuncSetAttribute(tkl_ts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dev.max_smem_optin));
      attr_set[1][di] = true;
    }
    MMB_CHECK_CUDA(launch_pdl(tkl_ts_kernel<1>));
  }
  MMB_CHECK_CUDA(cudaGetLastError());
#ifdef MMB200_ENABLE_PROF
  if (do_prof) {
    long long h[45 + 4 * 160];
    MMB_CHECK_CUDA(cudaStreamSynchronize(stream));
    MMB_CHECK_CUDA(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
    {
      long long mx[4] = {0, 0, 0, 0}, sm[4] = {0, 0, 0, 0};
      int arg = 0;
      for (int x = 0; x < grid && x < 160; ++x)
        for (int i = 0; i < 4; ++i) {
          const long long v = h[45 + 4 * x + i];
          sm[i] += v;
          if (v > mx[i]) { mx[i] = v; if (i == 0) arg = x; }
        }
      fprintf(stderr, "tkl_ts_prof per CTA (max / mean): epilogue total %lld / %lld (slowest CTA %d: phaseB %lld wait_accfull %lld) | phaseB %lld / %lld | "
              "wait_accfull %lld / %lld | mma role total %lld / %lld\n", mx[0], sm[0] / grid, arg, h[45 + 4 * arg + 1], h[45 + 4 * arg + 2], mx[1],
              sm[1] / grid, mx[2], sm[2] / grid, mx[3], sm[3] / grid);
    }
    MMB_CHECK_CUDA(cudaFree(prof));
    fprintf(stderr,
            "tkl_ts_prof cycles (CTA 0): tma total %lld wait_raw_empty %lld | mma total %lld wait_accempty %lld wait_op_full %lld | qconv total %lld "
            "wait_raw_full %lld wait_op_empty %lld | dconv total %lld wait_raw_full %lld wait_op_empty %lld | epi total %lld wait_accfull %lld "
            "phaseA %lld barriers23 %lld counts %lld phaseB %lld barrier4 %lld between %lld\n",
            h[8], h[0], h[17], h[9], h[10], h[26], h[18], h[19], h[35], h[27], h[28], h[44], h[36], h[37], h[38], h[39], h[40], h[41], h[43]);
  }
#endif
  *handled = true;
  return MMB200_OK;
}

}  // namespace mmb
