// Exact maximum-inner-product search with a fused per-query top-k (BERT_DOT dense retrieval scoring).
//
// Reference: `faiss.IndexIDMap(IndexFlatIP)` sharded over GPUs, called from
// matchmaker/retrieval/faiss_indices.py:27 (add_with_ids), :34 (search), :61-67 (shard=True, useFloat16),
// driven by matchmaker/dense_retrieval.py:328,391.  faiss tiles a cuBLAS GEMM into a score buffer and
// runs a k-selection kernel over it; here the [queries x passages] score matrix is never written:
//
//   flat_ip_tc_kernel   persistent CTAs over work items (block of 128 queries) x (range of passages); clusters of 2
//               CTAs take consecutive query blocks and share every passage tile by TMA multicast.
//       warp 0  TMA producer: per k-block one 16 KB query tile + this CTA's slice of the 32 KB passage tile
//       warp 1  tcgen05.mma issuer (whole-warp loop, elect.sync): D[128 queries x 256 passages] fp32 in TMEM, 2 slots
//       warps 2-9 epilogue: two warps per TMEM lane quarter, each filtering one 128-column half of every tile, thread =
//               query row: tcgen05.ld, FMNMX tree -> maxima of 8-column sub-groups, compare against the row's running
//               threshold tau (the k-th best seen so far); sub-groups holding a candidate for SOME row of the warp (~1/3
//               of them) are scanned with warp-uniform control flow and a predicated shared-memory atomic + global
//               store into the row's candidate list (ONE list per row, capacity 1024, global memory).  The pair of
//               warps meets at a named barrier at the start of every tile; rows whose list could overflow during the
//               tile are compacted there (split between the two warps): 32-step bisection on the order-preserving
//               integer image of the scores finds the k-th largest, survivors are rewritten in place and tau rises.
//               tau is also published per query (atomicMax) so items working on other passage ranges of the same
//               queries filter harder.  The per-tile loop must stay inside the instruction cache (compact_row is
//               __noinline__).
//   topk_merge_kernel   per query: bitonic sort of the candidate lists of all ranges (or, after the NCCL
//       all-gather, of all ranks) under the total order (score desc, id asc) -> [k] scores + ids.
//
// Tensor cores are used here because this is the one genuinely dense contraction of the hot path
// (arithmetic intensity ~ nq flops per passage byte).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "host_util.cuh"
#include "ptx.cuh"

namespace mmb {

namespace {

constexpr int kEpiPerQuarter = 2;        // epilogue warps per TMEM lane quarter (column halves of a tile, ONE list per row)
constexpr int kThreads = 64 + 128 * kEpiPerQuarter;  // TMA warp, MMA warp, 8 epilogue warps
// list entries per lane in a compaction: EPL = 32 -> capacity 1024 per row (k <= 256), EPL = 64 -> 2048 (k <= 1024)
constexpr int kMaxK = 1024;
__host__ __device__ constexpr int epl_for_k(int k) { return k <= 256 ? 32 : 64; }
constexpr int BM = 128;                 // queries per block (UMMA M)
constexpr int BN = 256;                 // passages per tile (UMMA N)
constexpr int kABytes = BM * 128;       // one k-block (64 halfs) of the query tile
constexpr int kBBytes = BN * 128;
constexpr int kStageBytes = kABytes + kBBytes;  // 48 KB
constexpr int kStages = 4;
constexpr int kAccSlots = 2;
constexpr int kMaxRanges = 32;

struct FipShared {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t accfull[kAccSlots];
  uint64_t accempty[kAccSlots];
  uint32_t tmem_base;
  uint32_t pad;
  int cnt[BM];        // entries in each row's candidate list (appended to by both warps of the row's quarter)
  uint32_t tau[BM];   // order-preserving image of each row's threshold
};

struct FipParams {
  const int64_t* ids;       // [n_pass] user ids or nullptr (id = id_base + position)
  int64_t id_base;
  int64_t nq, n_pass;
  int32_t dim, k, kpad;             // kpad = k rounded up to 32 (list capacity per row is 32 * EPL)
  int32_t kblocks, kb_wrap;         // k-blocks of 64 along the QUERY rows; passage k-block = kb < kb_wrap ? kb : kb - kb_wrap
  int32_t n_qblocks, n_ranges, tiles_per_range, n_tiles;
  int32_t fmt;
  uint2* lists;             // [grid][BM][cap]  (score bits, position)
  uint32_t* tau_glob;       // [nq] order-preserving image of the per-query threshold
  float* cand_scores;       // [nq][n_ranges * kpad]
  int64_t* cand_ids;        // [nq][n_ranges * kpad]
};

// order-preserving map float -> uint32 (larger float <=> larger key)
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr uint32_t kKeyNegInf = 0x007fffffu;  // f2key(-inf)

__device__ __forceinline__ int64_t pos_to_id(const FipParams& P, uint32_t pos) {
  return P.ids ? P.ids[pos] : P.id_base + (int64_t)pos;
}

// Warp-cooperative compaction of one row's candidate list to its top-k under (score desc, id asc).
// `list` has `cnt` valid entries (cnt <= 32 * EPL).  Returns the new count (min(cnt, k)) and the key of
// the k-th best entry in *kth_key (kKeyNegInf if fewer than k entries).
// __noinline__: the epilogue's per-tile loop has to stay inside the instruction cache.  With this routine inlined (and
// the column loop unrolled) the loop body streamed ~100 KB of code per tile and ran at IPC 0.03.
template <int EPL>
__device__ __noinline__ int compact_row(const FipParams& P, uint2* list, int cnt, int lane, uint32_t* kth_key) {
  uint32_t key[EPL], pos[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int e = lane + 32 * j;
    if (e < cnt) {
      const uint2 v = list[e];
      key[j] = f2key(__uint_as_float(v.x));
      pos[j] = v.y;
    } else {
      key[j] = 0u;  // below every real key (real keys are >= kKeyNegInf > 0 unless NaN; NaNs are not supported)
      pos[j] = 0xffffffffu;
    }
  }
  if (cnt <= P.k) {
    *kth_key = cnt == P.k ? 0u : kKeyNegInf;
    if (cnt == P.k) {  // exactly k: threshold = smallest key present
      uint32_t mn = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (lane + 32 * j < cnt) mn = min(mn, key[j]);
      *kth_key = __reduce_min_sync(0xffffffffu, mn);
    }
    return cnt;
  }
  // largest T with count(key >= T) >= k
  uint32_t lo = 0u, hi = 0xffffffffu;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;  // upper mid
    int c = 0;
#pragma unroll
    for (int j = 0; j < EPL; ++j) c += (key[j] >= mid) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= P.k) lo = mid; else hi = mid - 1u;
  }
  const uint32_t T = lo;
  int c_gt = 0, c_eq = 0;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    c_gt += (key[j] > T) ? 1 : 0;
    c_eq += (key[j] == T) ? 1 : 0;
  }
  c_gt = __reduce_add_sync(0xffffffffu, c_gt);
  c_eq = __reduce_add_sync(0xffffffffu, c_eq);
  int need = P.k - c_gt;  // how many of the entries tied at T survive (1 <= need <= c_eq)
  uint64_t keep = 0u;     // bit j: entry j of this lane survives
#pragma unroll
  for (int j = 0; j < EPL; ++j)
    if (key[j] > T) keep |= 1ull << j;
  if (need == c_eq) {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (key[j] == T) keep |= 1ull << j;
  } else {
    // rare: more ties than room -> take the `need` smallest ids among them
    uint64_t taken = 0u;
    for (int n = 0; n < need; ++n) {
      unsigned long long best = ~0ull;
      int bj = -1;
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (key[j] == T && !(taken & (1ull << j))) {
          const unsigned long long id = (unsigned long long)(pos_to_id(P, pos[j]) ^ (1ll << 63));  // signed order
          if (id < best) { best = id; bj = j; }
        }
      unsigned long long wbest = best;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ot = __shfl_xor_sync(0xffffffffu, wbest, o);
        wbest = ot < wbest ? ot : wbest;
      }
      const unsigned owner = __ballot_sync(0xffffffffu, best == wbest && bj >= 0);
      if (bj >= 0 && best == wbest && (int)(__ffs(owner) - 1) == lane) { taken |= 1ull << bj; keep |= 1ull << bj; }
    }
  }
  int base = 0;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const bool kp = (keep >> j) & 1ull;
    const unsigned b = __ballot_sync(0xffffffffu, kp);
    if (kp) list[base + __popc(b & ((1u << lane) - 1u))] = make_uint2(__float_as_uint(key2f(key[j])), pos[j]);
    base += __popc(b);
  }
  __syncwarp();
  *kth_key = T;
  return P.k;
}

// CL = thread-block cluster size.  The CL CTAs of a cluster work on CL consecutive query blocks against the SAME passage
// tiles: each CTA fetches 1/CL of every passage tile and multicasts it to the whole cluster, so the L2 -> SM traffic per
// CTA drops from 48 KB to (16 + 32 / CL) KB per k-block (7.2 TB/s of L2 reads at CL = 1).  A stage may be refilled only
// when EVERY CTA of the cluster has consumed it, hence the multicast commit onto all `empty` barriers (count CL).
// PROF (MMB200_FLATIP_PROF=1): debugging aid, one thread per role of CTA 0 accumulates the cycles it spends blocked
#define FIP_TIMED(slot, stmt)                                 \
  do {                                                        \
    if constexpr (PROF) {                                     \
      const long long t0_ = clock64();                        \
      stmt;                                                   \
      pc[slot] += clock64() - t0_;                            \
    } else {                                                  \
      stmt;                                                   \
    }                                                         \
  } while (0)

template <int CL, bool PROF = false, int EPL = 32>
__global__ void __launch_bounds__(kThreads, 1)
flat_ip_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_p, FipParams P,
                  long long* prof = nullptr) {
  long long pc[3] = {0, 0, 0};
  const long long t_start = PROF ? clock64() : 0;
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment for SWIZZLE_128B tiles, derived by pointer arithmetic on the __shared__ array so the
  // compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  FipShared* S = reinterpret_cast<FipShared*>(smem + (size_t)kStages * kStageBytes);
  constexpr int kCap = 32 * EPL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = P.kblocks;
  const int n_qgroups = (P.n_qblocks + CL - 1) / CL;   // CL consecutive query blocks per cluster work item
  const int n_items = n_qgroups * P.n_ranges;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CL) - 1u);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_p);
    for (int s = 0; s < kStages; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], CL); }
    for (int s = 0; s < kAccSlots; ++s) { mbar_init(&S->accfull[s], 1); mbar_init(&S->accempty[s], 4 * kEpiPerQuarter); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&S->tmem_base, 512);
  tc_fence_before_sync();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // peers signal our barriers: their init must be visible cluster-wide
  tc_fence_after_sync();
  const uint32_t tmem_base = S->tmem_base;

  if (warp == 0) {
    // the whole warp walks the loop (uniform control flow and operands); one elected lane issues the TMA: inside an
    // `if (lane == 0)` region the compiler wraps every TMA / MMA in an ELECT / R2UR waterfall loop (ptx.cuh)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int rg = item / n_qgroups, qb = (item % n_qgroups) * CL + rank;  // range-major: co-running CTAs share passages
        const int t0 = rg * P.tiles_per_range, t1 = min(P.n_tiles, t0 + P.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < kblocks; ++kb) {
            FIP_TIMED(0, mbar_wait(&S->empty[stage], phase ^ 1u));
            uint8_t* st = smem + (size_t)stage * kStageBytes;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&S->full[stage], (uint32_t)kStageBytes);
              const int kbp = kb < P.kb_wrap ? kb : kb - P.kb_wrap;   // fp32-split storage: [q_hi|q_lo|q_hi] x [p_hi|p_hi|p_lo]
              tma_load_2d(&tmap_q, st, &S->full[stage], kb * 64, qb * BM, kEvictLast);
              if (CL == 1)
                tma_load_2d(&tmap_p, st + kABytes, &S->full[stage], kbp * 64, t * BN, kEvictFirst);
              else  // this CTA's slice of the passage tile, written into every CTA of the cluster
                tma_load_2d_multicast(&tmap_p, st + kABytes + rank * (kBBytes / CL), &S->full[stage], kbp * 64,
                                      t * BN + rank * (BN / CL), kAllCtas, kEvictFirst);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = make_idesc((uint32_t)P.fmt, BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, accphase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int rg = item / n_qgroups;
        const int t0 = rg * P.tiles_per_range, t1 = min(P.n_tiles, t0 + P.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          FIP_TIMED(0, mbar_wait(&S->accempty[acc], accphase ^ 1u));
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < kblocks; ++kb) {
            FIP_TIMED(1, mbar_wait(&S->full[stage], phase));
            tc_fence_after_sync();
            const long long t_i = PROF ? clock64() : 0;
            const uint32_t a = smem_u32(smem + (size_t)stage * kStageBytes);
            const uint64_t da = make_sw128_kmajor_desc(a), db = make_sw128_kmajor_desc(a + kABytes);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)  // +32 bytes along K inside the 128-byte swizzle atom = +2 in the address field
                umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
              if (CL == 1) umma_commit(&S->empty[stage]); else umma_commit_multicast(&S->empty[stage], kAllCtas);
              if (kb == kblocks - 1) umma_commit(&S->accfull[acc]);
            }
            __syncwarp();
            if (PROF) pc[2] += clock64() - t_i;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (++acc == kAccSlots) { acc = 0; accphase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------- epilogue: filter + top-k lists ---------------------------------
    // Two warps per TMEM lane quarter (one epilogue warp per scheduler runs its dependent chains at ~0.2 IPC and cannot
    // keep up with the tensor pipe): warp (quarter, half) filters columns [128 half, 128 half + 128) of every tile.  Both
    // append to the SAME per-row list (shared-memory counter, atomicAdd) under the SAME threshold, so nothing about the
    // selection changes.  The pair meets at a named barrier at the start of every tile: counters are final there, both
    // warps see the same set of nearly-full rows and split their compaction.  A row gains at most 256 entries per
    // tile, so compacting when cnt > cap - 256 keeps every append in bounds.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;  // query row inside the block == TMEM lane
    const uint32_t pair_bar = 1u + (uint32_t)quarter;
    int acc = 0;
    uint32_t accphase = 0;
    uint2* my_list = P.lists + ((size_t)blockIdx.x * BM + row) * kCap;
    uint2* warp_lists = P.lists + ((size_t)blockIdx.x * BM + quarter * 32) * kCap;
    int* cnt_s = S->cnt + quarter * 32;
    uint32_t* tau_s = S->tau + quarter * 32;
    const uint32_t my_cnt_addr = smem_u32(cnt_s + lane);
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int rg = item / n_qgroups, qb = (item % n_qgroups) * CL + rank;  // range-major: co-running CTAs share passages
      const int t0 = rg * P.tiles_per_range, t1 = min(P.n_tiles, t0 + P.tiles_per_range);
      const int64_t q = (int64_t)qb * BM + row;
      const bool live = q < P.nq;
      uint32_t tau_seen = live ? P.tau_glob[q] : 0xffffffffu;  // dead rows accept nothing
      if (half == 0) { cnt_s[lane] = 0; tau_s[lane] = tau_seen; }
      for (int t = t0; t < t1; ++t) {
        FIP_TIMED(0, mbar_wait(&S->accfull[acc], accphase));
        tc_fence_after_sync();
        const long long t_e = PROF ? clock64() : 0;
        if (half == 0 && live) tau_s[lane] = max(tau_s[lane], tau_seen);   // other ranges' progress, fetched a tile ago
        named_bar_sync(pair_bar, 64);      // appends of the previous tile are complete, counters and thresholds final
        if (live) tau_seen = *reinterpret_cast<volatile const uint32_t*>(P.tau_glob + q);  // consumed one tile later
        const unsigned full_rows = __ballot_sync(0xffffffffu, cnt_s[lane] > kCap - BN);
        named_bar_sync(pair_bar, 64);      // both warps have read the counters before either appends to them again
        if (full_rows) {  // same value in both warps: rows are dealt alternately
          int idx = 0;
          for (unsigned m = full_rows; m; m &= m - 1, ++idx) {
            if ((idx & 1) != half) continue;
            const int rr = __ffs(m) - 1;
            uint32_t kth;
            const int nc = compact_row<EPL>(P, warp_lists + (size_t)rr * kCap, cnt_s[rr], lane, &kth);
            __syncwarp();
            if (lane == 0) {
              cnt_s[rr] = nc;
              tau_s[rr] = max(tau_s[rr], kth);
              const int64_t qq = (int64_t)qb * BM + quarter * 32 + rr;
              if (qq < P.nq) atomicMax(P.tau_glob + qq, kth);
            }
          }
          named_bar_sync(pair_bar, 64);
        }
        const float tau = key2f(tau_s[lane]);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
        const int64_t col0 = (int64_t)t * BN + half * (BN / 2);
        const bool ragged = (int64_t)t * BN + BN > P.n_pass;
        // all 128 columns of this warp into registers, then the accumulator slot goes straight back to the MMA warp:
        // with two slots the tensor pipe otherwise idles while the filter below works its way through the tile
        uint32_t r4[BN / 2 / 32][32];
#pragma unroll
        for (int c = 0; c < BN / 2 / 32; ++c) tmem_ld_32x32b_x32(taddr + c * 32, r4[c]);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->accempty[acc]);
#pragma unroll
        for (int c = 0; c < BN / 2 / 32; ++c) {
          uint32_t (&r)[32] = r4[c];
          const uint32_t pbase = (uint32_t)(col0 + c * 32);
          // A row sees a candidate in a few % of its 32-column groups, but SOME row of the warp does in most of them,
          // so the path behind the maxima has to be cheap for the idle lanes too: four sub-groups of 8 whose maxima
          // come out of one FMNMX tree, votes issued back to back, and only sub-groups with a candidate are scanned --
          // under warp-uniform control flow (per-lane branches cost a BSSY / BSYNC pair per element) with a
          // predicated atomic + store.
          float g4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float m0 = fmaxf(fmaxf(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])), __uint_as_float(r[8 * i + 2]));
            const float m1 = fmaxf(fmaxf(__uint_as_float(r[8 * i + 3]), __uint_as_float(r[8 * i + 4])), __uint_as_float(r[8 * i + 5]));
            g4[i] = fmaxf(fmaxf(m0, m1), fmaxf(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
          }
          const unsigned bal[4] = {__ballot_sync(0xffffffffu, g4[0] >= tau), __ballot_sync(0xffffffffu, g4[1] >= tau),
                                   __ballot_sync(0xffffffffu, g4[2] >= tau), __ballot_sync(0xffffffffu, g4[3] >= tau)};
          if ((bal[0] | bal[1]) | (bal[2] | bal[3])) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (bal[i]) {
                uint32_t e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  bool pass = __uint_as_float(r[8 * i + j]) >= tau;
                  if (ragged) pass = pass && (int64_t)(pbase + 8 * i + j) < P.n_pass;
                  e[j] = pass ? (1u << j) : 0u;
                }
                uint32_t m = ((e[0] | e[1]) | (e[2] | e[3])) | ((e[4] | e[5]) | (e[6] | e[7]));
                while (__any_sync(0xffffffffu, m != 0)) {
                  const int j = (__ffs(m) - 1) & 7;  // 7 for lanes that are done (nothing stored)
                  const uint32_t a01 = (j & 1) ? r[8 * i + 1] : r[8 * i + 0], a23 = (j & 1) ? r[8 * i + 3] : r[8 * i + 2];
                  const uint32_t a45 = (j & 1) ? r[8 * i + 5] : r[8 * i + 4], a67 = (j & 1) ? r[8 * i + 7] : r[8 * i + 6];
                  const uint32_t a03 = (j & 2) ? a23 : a01, a47 = (j & 2) ? a67 : a45;
                  const uint32_t v = (j & 4) ? a47 : a03;
                  uint32_t slot;
                  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmov.u32 %0, 0;\n\t@p atom.shared.add.u32 %0, [%2], 1;\n\t}"
                               : "=r"(slot)
                               : "r"(m), "r"(my_cnt_addr)
                               : "memory");
                  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.v2.u32 [%1], {%2, %3};\n\t}"
                               ::"r"(m), "l"(my_list + slot), "r"(v), "r"(pbase + 8 * i + j)
                               : "memory");
                  m &= m - 1;
                }
              }
            }
          }
        }
        if (PROF) pc[1] += clock64() - t_e;
        if (++acc == kAccSlots) { acc = 0; accphase ^= 1u; }
      }
      // item done: final compaction of every row (split between the two warps), then publish (score, id) candidates
      const long long t_f = PROF ? clock64() : 0;
      named_bar_sync(pair_bar, 64);
      for (int rr = half; rr < 32; rr += 2) {
        uint32_t kth;
        const int nc = compact_row<EPL>(P, warp_lists + (size_t)rr * kCap, cnt_s[rr], lane, &kth);
        const int64_t qq = (int64_t)qb * BM + quarter * 32 + rr;
        if (qq < P.nq) {
          if (lane == 0 && nc == P.k) atomicMax(P.tau_glob + qq, kth);
          const uint2* lst = warp_lists + (size_t)rr * kCap;
          float* cs = P.cand_scores + (size_t)qq * P.n_ranges * P.kpad + (size_t)rg * P.kpad;
          int64_t* ci = P.cand_ids + (size_t)qq * P.n_ranges * P.kpad + (size_t)rg * P.kpad;
          for (int e = lane; e < P.kpad; e += 32) {
            if (e < nc) {
              const uint2 v = lst[e];
              cs[e] = __uint_as_float(v.x);
              ci[e] = pos_to_id(P, v.y);
            } else {
              cs[e] = -INFINITY;
              ci[e] = -1;
            }
          }
        }
        __syncwarp();
      }
      named_bar_sync(pair_bar, 64);   // the counters are reset by the next item only after both warps are done with them
      if (PROF) pc[2] += clock64() - t_f;
    }
    if (PROF && blockIdx.x == 0 && threadIdx.x == 64) { prof[5] = pc[0]; prof[6] = pc[1]; prof[7] = pc[2]; }
  }
  if (PROF && blockIdx.x == 0 && lane == 0) {
    if (warp == 0) prof[0] = pc[0];
    if (warp == 1) { prof[1] = pc[0]; prof[2] = pc[1]; prof[3] = pc[2]; }
  }

  tc_fence_before_sync();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA may exit while peers still multicast into it
  if (PROF && blockIdx.x == 0 && threadIdx.x == 0) prof[4] = clock64() - t_start;
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}
#undef FIP_TIMED

// ---------------------------------------------------------------------------------------------
// merge: per query, sort L candidates by (score desc, id asc), emit the first k.
// ---------------------------------------------------------------------------------------------
struct Cand {
  float s;
  int32_t valid;
  int64_t id;
};
__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {
  if (a.valid != b.valid) return a.valid > b.valid;
  if (a.s != b.s) return a.s > b.s;
  return a.id < b.id;
}

// Block (q, grp) sorts candidates [grp * seg, min(L, (grp + 1) * seg)) of query q and writes its k best to row
// q * n_groups + grp of the output.  A candidate is void when its score is NaN, -inf or faiss's "no result" value
// (-FLT_MAX) -- NOT by the sign of its id: faiss IndexIDMap accepts negative user ids.  `final_pass` selects the
// filler for missing results: (-FLT_MAX, -1) as faiss returns them, or (-inf, -1) between passes.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ cand_scores,
                                                         const int64_t* __restrict__ cand_ids, int64_t nq, int L, int seg,
                                                         int n_groups, int Lpow2, int k, int final_pass,
                                                         float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ __align__(16) uint8_t msm[];
  Cand* c = reinterpret_cast<Cand*>(msm);
  for (int64_t item = blockIdx.x; item < nq * n_groups; item += gridDim.x) {
    const int64_t q = item / n_groups;
    const int grp = (int)(item % n_groups);
    const int lo = grp * seg, n = min(seg, L - lo);
    __syncthreads();
    for (int e = threadIdx.x; e < Lpow2; e += blockDim.x) {
      Cand v;
      if (e < n) {
        v.s = cand_scores[q * L + lo + e];
        v.id = cand_ids[q * L + lo + e];
        v.valid = (v.s == v.s && v.s > -3.4028234663852886e38f) ? 1 : 0;
      } else {
        v.s = -INFINITY; v.id = -1; v.valid = 0;
      }
      c[e] = v;
    }
    __syncthreads();
    for (int size = 2; size <= Lpow2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int e = threadIdx.x; e < Lpow2 / 2; e += blockDim.x) {
          const int i = 2 * e - (e & (stride - 1));
          const int j = i + stride;
          const bool up = ((i & size) == 0);
          const Cand a = c[i], b = c[j];
          const bool swap = up ? cand_before(b, a) : cand_before(a, b);
          if (swap) { c[i] = b; c[j] = a; }
        }
        __syncthreads();
      }
    }
    for (int e = threadIdx.x; e < k; e += blockDim.x) {
      const bool ok = e < Lpow2 && c[e].valid;
      out_scores[item * k + e] = ok ? c[e].s : (final_pass ? -3.4028234663852886e38f : -INFINITY);
      out_ids[item * k + e] = ok ? c[e].id : -1;
    }
  }
}

struct Plan {
  int n_qblocks, n_tiles, n_ranges, tiles_per_range, grid, kpad, cl;
};

Plan make_plan(int64_t nq, int64_t n_pass, int k, int sm_count) {
  Plan pl;
  pl.n_qblocks = (int)((nq + BM - 1) / BM);
  pl.n_tiles = (int)((n_pass + BN - 1) / BN);
  pl.kpad = (k + 31) / 32 * 32;
  // cluster size: CTAs of a cluster take consecutive query blocks and share every passage tile by TMA multicast.  A
  // cluster slot without a query block still runs (its slice of the passage tile is needed by its peers), so only pair
  // up when little is wasted.  MMB200_FLATIP_CLUSTER overrides (1, 2 or 4).
  pl.cl = (pl.n_qblocks % 2 == 0 || pl.n_qblocks >= 9) ? 2 : 1;
  if (const char* env = getenv("MMB200_FLATIP_CLUSTER")) {
    const int c = atoi(env);
    if (c == 1 || c == 2 || c == 4) pl.cl = c;
  }
  const int n_qgroups = (pl.n_qblocks + pl.cl - 1) / pl.cl;
  const int max_clusters = std::max(1, sm_count / pl.cl);
  // number of passage ranges: every range restarts its threshold at -inf and pays ~log(range/k) list
  // compactions per query, so take the SMALLEST count whose item grid fills the SMs to >= 88 % (or the best
  // fill available).  MMB200_FLATIP_RANGES overrides it for experiments.
  int best_r = 1;
  double best_eff = -1.0;
  const int max_r = std::max(1, std::min(kMaxRanges, pl.n_tiles));
  double effs[kMaxRanges + 1];
  for (int r = 1; r <= max_r; ++r) {
    const int64_t items = (int64_t)n_qgroups * r;
    const int64_t g = std::min<int64_t>(max_clusters, items);
    const int64_t waves = (items + g - 1) / g;
    effs[r] = (double)items / (double)(waves * max_clusters);
    if (effs[r] > best_eff + 1e-9) { best_eff = effs[r]; best_r = r; }
  }
  for (int r = 1; r <= max_r; ++r)
    if (effs[r] >= 0.88 || effs[r] >= best_eff - 1e-9) { best_r = r; break; }
  if (const char* env = getenv("MMB200_FLATIP_RANGES")) {
    const int r = atoi(env);
    if (r >= 1 && r <= max_r) best_r = r;
  }
  pl.tiles_per_range = (pl.n_tiles + best_r - 1) / best_r;
  pl.n_ranges = (pl.n_tiles + pl.tiles_per_range - 1) / pl.tiles_per_range;
  pl.grid = pl.cl * (int)std::min<int64_t>(max_clusters, (int64_t)n_qgroups * pl.n_ranges);
  return pl;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t workspace_bytes(const Plan& pl, int64_t nq, int k) {
  const size_t cap = 32 * (size_t)epl_for_k(k);
  return align256((size_t)nq * sizeof(uint32_t)) + align256((size_t)pl.grid * BM * cap * sizeof(uint2)) +
         align256((size_t)nq * pl.n_ranges * pl.kpad * sizeof(float)) +
         align256((size_t)nq * pl.n_ranges * pl.kpad * sizeof(int64_t));
}

size_t total_workspace_bytes(int64_t nq, int64_t n_pass, int k, int sm_count) {
  return workspace_bytes(make_plan(nq, n_pass, k, sm_count), nq, k);
}

__global__ void fill_u32(uint32_t* p, int64_t n, uint32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Largest candidate count one shared-memory sort takes (16 B per candidate).  More than that -- many passage ranges at
// k = 1024, or sharding.topk_all_gather_merge over a shard with thousands of local scores per query -- is merged in
// passes: groups of kMergeSeg candidates are cut to their k best, the survivors merged again.
constexpr int kMergeSeg = 8192;

int launch_merge(const float* cand_scores, const int64_t* cand_ids, int64_t nq, int L, int k, float* out_scores,
                 int64_t* out_ids, const DeviceInfo& dev, cudaStream_t stream) {
  MMB_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(kMergeSeg * sizeof(Cand))));
  const float* in_s = cand_scores;
  const int64_t* in_i = cand_ids;
  float* tmp_s[2] = {nullptr, nullptr};
  int64_t* tmp_i[2] = {nullptr, nullptr};
  int cur = 0;
  int rc = MMB200_OK;
  while (true) {
    const bool last = L <= kMergeSeg;
    const int seg = last ? L : kMergeSeg;
    const int groups = (L + seg - 1) / seg;
    int lp = 2;
    while (lp < seg) lp <<= 1;
    const int k_out = last ? k : std::min(k, seg);
    float* o_s = out_scores;
    int64_t* o_i = out_ids;
    if (!last) {
      if (cudaMallocAsync(reinterpret_cast<void**>(&tmp_s[cur]), (size_t)nq * groups * k_out * sizeof(float), stream) != cudaSuccess ||
          cudaMallocAsync(reinterpret_cast<void**>(&tmp_i[cur]), (size_t)nq * groups * k_out * sizeof(int64_t), stream) != cudaSuccess) {
        set_error("topk merge: cannot allocate the intermediate candidate lists");
        rc = MMB200_ERR_CUDA;
        break;
      }
      o_s = tmp_s[cur];
      o_i = tmp_i[cur];
    }
    const int grid = (int)std::min<int64_t>(nq * groups, (int64_t)dev.sm_count * 4);
    topk_merge_kernel<<<grid, 256, (size_t)lp * sizeof(Cand), stream>>>(in_s, in_i, nq, L, seg, groups, lp, k_out, last ? 1 : 0,
                                                                        o_s, o_i);
    if (cudaGetLastError() != cudaSuccess) {
      set_error("topk merge: kernel launch failed");
      rc = MMB200_ERR_CUDA;
      break;
    }
    if (last) break;
    in_s = o_s;
    in_i = o_i;
    L = groups * k_out;
    cur ^= 1;
    if (tmp_s[cur]) {  // the buffers of two passes ago are no longer read by anything enqueued after this point
      cudaFreeAsync(tmp_s[cur], stream);
      cudaFreeAsync(tmp_i[cur], stream);
      tmp_s[cur] = nullptr;
      tmp_i[cur] = nullptr;
    }
  }
  for (int i = 0; i < 2; ++i) {
    if (tmp_s[i]) cudaFreeAsync(tmp_s[i], stream);
    if (tmp_i[i]) cudaFreeAsync(tmp_i[i], stream);
  }
  return rc;
}

}  // namespace

}  // namespace mmb

extern "C" int64_t mmb200_flat_ip_workspace_bytes(int64_t nq, int64_t n_pass, int32_t k) {
  using namespace mmb;
  if (nq <= 0 || n_pass <= 0 || k <= 0 || k > kMaxK) return 0;
  DeviceInfo dev;
  if (current_device_info(&dev)) return -1;
  return (int64_t)total_workspace_bytes(nq, n_pass, k, dev.sm_count);
}

extern "C" int mmb200_flat_ip_plan(int64_t nq, int64_t n_pass, int32_t k, int32_t sm_count, int32_t out[8]) {
  using namespace mmb;
  MMB_REQUIRE(out != nullptr, "null pointer");
  MMB_REQUIRE(nq > 0 && n_pass > 0 && k >= 1 && k <= kMaxK && sm_count >= 1, "bad sizes");
  MMB_REQUIRE(n_pass < (1ll << 32) - 512, "at most 2^32 passages per shard");
  const Plan pl = make_plan(nq, n_pass, k, sm_count);
  const uint64_t ws = (uint64_t)workspace_bytes(pl, nq, k);
  out[0] = pl.n_qblocks; out[1] = pl.n_tiles; out[2] = pl.n_ranges; out[3] = pl.tiles_per_range; out[4] = pl.grid; out[5] = pl.cl;
  out[6] = (int32_t)(uint32_t)(ws & 0xffffffffu); out[7] = (int32_t)(uint32_t)(ws >> 32);
  return MMB200_OK;
}

extern "C" int mmb200_flat_ip_topk(const void* queries, const void* passages, const int64_t* ids, float* out_scores,
                                   int64_t* out_ids, void* workspace, int64_t workspace_bytes_given, int64_t nq,
                                   int64_t n_pass, int32_t dim, int32_t k, int32_t dtype, int64_t id_base,
                                   void* stream_) {
  using namespace mmb;
  MMB_REQUIRE(queries && passages && out_scores && out_ids && workspace, "null pointer");
  MMB_REQUIRE(nq > 0 && n_pass > 0, "need at least one query and one passage");
  MMB_REQUIRE(k >= 1 && k <= kMaxK, "fused top-k supports 1 <= k <= 1024");
  MMB_REQUIRE(dtype == MMB200_F16 || dtype == MMB200_BF16 || dtype == MMB200_F32_SPLIT16,
              "passage storage must be fp16, bf16 or the fp16 hi/lo split of fp32 (MMB200_F32_SPLIT16)");
  const bool split = dtype == MMB200_F32_SPLIT16;
  MMB_REQUIRE(dim % 64 == 0 && dim >= 64, "vector dim must be a multiple of 64");
  MMB_REQUIRE(n_pass < (1ll << 32) - 512, "at most 2^32 passages per shard");
  MMB_REQUIRE(((reinterpret_cast<uintptr_t>(queries) | reinterpret_cast<uintptr_t>(passages)) & 15) == 0, "16-byte alignment");
  DeviceInfo dev;
  if (int rc = current_device_info(&dev)) return rc;
  if (!is_sm100(dev)) {
    set_error("matchmaker_b200 kernels are built for sm_100a only");
    return MMB200_ERR_UNSUPPORTED;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Plan pl = make_plan(nq, n_pass, k, dev.sm_count);
  MMB_REQUIRE((size_t)workspace_bytes_given >= total_workspace_bytes(nq, n_pass, k, dev.sm_count),
              "workspace too small (see mmb200_flat_ip_workspace_bytes)");
  const CUtensorMapDataType tdt = dtype == MMB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const uint64_t q_cols = split ? 3ull * dim : (uint64_t)dim, p_cols = split ? 2ull * dim : (uint64_t)dim;
  CUtensorMap tq;
  {
    const uint64_t dims[2] = {q_cols, (uint64_t)nq};
    const uint64_t strides[1] = {q_cols * 2};
    const uint32_t box[2] = {64, BM};
    if (int rc = encode_tensor_map(&tq, tdt, 2, queries, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  uint32_t* tau_glob = static_cast<uint32_t*>(workspace);
  fill_u32<<<64, 256, 0, stream>>>(tau_glob, nq, kKeyNegInf);
  MMB_CHECK_CUDA(cudaGetLastError());

  // one pass of flat_ip_tc_kernel over `n_rows` passages whose rows are `row_pitch` bytes apart
  auto run_pass = [&](const Plan& pp, int64_t n_rows, uint64_t row_pitch, const int64_t* pass_ids, int64_t pass_id_base,
                      FipParams* out_params) -> int {
    FipParams P{};
    uint8_t* w = static_cast<uint8_t*>(workspace) + align256((size_t)nq * sizeof(uint32_t));
    P.tau_glob = tau_glob;
    P.lists = reinterpret_cast<uint2*>(w);
    w += align256((size_t)pp.grid * BM * 32 * (size_t)epl_for_k(k) * sizeof(uint2));
    P.cand_scores = reinterpret_cast<float*>(w);
    w += align256((size_t)nq * pp.n_ranges * pp.kpad * sizeof(float));
    P.cand_ids = reinterpret_cast<int64_t*>(w);
    P.ids = pass_ids; P.id_base = pass_id_base; P.nq = nq; P.n_pass = n_rows; P.dim = dim; P.k = k; P.kpad = pp.kpad;
    P.n_qblocks = pp.n_qblocks; P.n_ranges = pp.n_ranges; P.tiles_per_range = pp.tiles_per_range; P.n_tiles = pp.n_tiles;
    P.fmt = dtype == MMB200_BF16 ? kFmtBF16 : kFmtF16;
    P.kblocks = (int32_t)(q_cols / 64);
    P.kb_wrap = split ? dim / 64 : P.kblocks;
    CUtensorMap tp;
    {
      const uint64_t dims[2] = {p_cols, (uint64_t)n_rows};
      const uint64_t strides[1] = {row_pitch};
      const uint32_t box[2] = {64, (uint32_t)(BN / pp.cl)};   // each CTA of a cluster fetches (and multicasts) its slice
      if (int rc = encode_tensor_map(&tp, tdt, 2, passages, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
        return rc;
    }
    const size_t smem = (size_t)kStages * kStageBytes + sizeof(FipShared) + 1024;
    auto launch = [&](auto kernel) -> int {
      MMB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)pp.grid);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = (unsigned)pp.cl;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      MMB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tq, tp, P, (long long*)nullptr));
      return MMB200_OK;
    };
    if (out_params) *out_params = P;
#ifdef MMB200_ENABLE_PROF
    if (pp.cl == 1 && out_params && epl_for_k(k) == 32 && getenv("MMB200_FLATIP_PROF")) {  // debugging aid: per-role wait cycles
      long long* prof = nullptr;
      long long h[12] = {0};
      MMB_CHECK_CUDA(cudaMalloc(&prof, sizeof(h)));
      MMB_CHECK_CUDA(cudaMemset(prof, 0, sizeof(h)));
      MMB_CHECK_CUDA(cudaFuncSetAttribute(flat_ip_tc_kernel<1, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      flat_ip_tc_kernel<1, true, 32><<<pp.grid, kThreads, smem, stream>>>(tq, tp, P, prof);
      MMB_CHECK_CUDA(cudaStreamSynchronize(stream));
      MMB_CHECK_CUDA(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
      MMB_CHECK_CUDA(cudaFree(prof));
      fprintf(stderr,
              "fip_prof cycles: total %lld | tma wait_empty %lld | mma wait_accempty %lld wait_full %lld issue %lld | "
              "epi wait_accfull %lld tile %lld item_end %lld\n",
              h[4], h[0], h[1], h[2], h[3], h[5], h[6], h[7]);
      return MMB200_OK;
    }
#endif
    if (epl_for_k(k) == 32)
      return pp.cl == 1   ? launch(flat_ip_tc_kernel<1, false, 32>)
             : pp.cl == 2 ? launch(flat_ip_tc_kernel<2, false, 32>)
                          : launch(flat_ip_tc_kernel<4, false, 32>);
    return pp.cl == 1   ? launch(flat_ip_tc_kernel<1, false, 64>)
           : pp.cl == 2 ? launch(flat_ip_tc_kernel<2, false, 64>)
                        : launch(flat_ip_tc_kernel<4, false, 64>);
  };

  FipParams P{};
  if (int rc = run_pass(pl, n_pass, p_cols * 2, ids, id_base, &P)) return rc;
  return launch_merge(P.cand_scores, P.cand_ids, nq, pl.n_ranges * pl.kpad, k, out_scores, out_ids, dev, stream);
}

extern "C" int mmb200_topk_merge(const float* cand_scores, const int64_t* cand_ids, float* out_scores, int64_t* out_ids,
                                 int64_t nq, int32_t n_candidates, int32_t k, void* stream_) {
  using namespace mmb;
  MMB_REQUIRE(cand_scores && cand_ids && out_scores && out_ids, "null pointer");
  MMB_REQUIRE(nq >= 0 && n_candidates >= 1 && k >= 1, "bad sizes");
  if (nq == 0) return MMB200_OK;
  DeviceInfo dev;
  if (int rc = current_device_info(&dev)) return rc;
  if (!is_sm100(dev)) {
    set_error("matchmaker_b200 kernels are built for sm_100a only");
    return MMB200_ERR_UNSUPPORTED;
  }
  return launch_merge(cand_scores, cand_ids, nq, n_candidates, k, out_scores, out_ids, dev, static_cast<cudaStream_t>(stream_));
}
