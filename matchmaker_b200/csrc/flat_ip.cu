// Exact maximum-inner-product search with a fused per-query top-k (BERT_DOT dense retrieval scoring).
//
// Reference: `faiss.IndexIDMap(IndexFlatIP)` sharded over GPUs, called from
// matchmaker/retrieval/faiss_indices.py:27 (add_with_ids), :34 (search), :61-67 (shard=True, useFloat16),
// driven by matchmaker/dense_retrieval.py:328,391.  faiss tiles a cuBLAS GEMM into a score buffer and
// runs a k-selection kernel over it; here the [queries x passages] score matrix is never written:
//
//   flat_ip_tc_kernel   persistent CTAs over work items (block of 128 queries) x (range of passages).
//       warp 0  TMA producer: per k-block one 16 KB query tile + one 32 KB passage tile (SWIZZLE_128B)
//       warp 1  tcgen05.mma issuer: D[128 queries x 256 passages] fp32 in TMEM, 2 accumulator slots
//       warps 2-5 epilogue: thread = query row; tcgen05.ld 32 columns at a time, compare against the row's
//               running threshold tau (the k-th best seen so far), append survivors to the row's private
//               candidate list (global memory, L2 resident), and when a list fills up the warp compacts
//               it cooperatively: 32-step bisection on the order-preserving integer image of the scores
//               finds the k-th largest, survivors are rewritten in place and tau rises.  tau is also
//               published per query (atomicMax) so items working on other passage ranges of the same
//               queries filter harder.  Expected appends per row ~ k * ln(n / k): the epilogue costs a few
//               hundred instructions per 256-column tile against 6144 MMA cycles.
//   topk_merge_kernel   per query: bitonic sort of the candidate lists of all ranges (or, after the NCCL
//       all-gather, of all ranks) under the total order (score desc, id asc) -> [k] scores + ids.
//
// Tensor cores are used here because this is the one genuinely dense contraction of the hot path
// (arithmetic intensity ~ nq flops per passage byte).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "host_util.cuh"
#include "ptx.cuh"

namespace mmb {

namespace {

constexpr int kThreads = 192;
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
};

struct FipParams {
  const int64_t* ids;       // [n_pass] user ids or nullptr (id = id_base + position)
  int64_t id_base;
  int64_t nq, n_pass;
  int32_t dim, k, kpad, cap;        // kpad = k rounded up to 32; cap = list capacity per row (32 * EPL)
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
template <int EPL>
__device__ __forceinline__ int compact_row(const FipParams& P, uint2* list, int cnt, int lane, uint32_t* kth_key) {
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
  uint32_t keep = 0u;     // bit j: entry j of this lane survives
#pragma unroll
  for (int j = 0; j < EPL; ++j)
    if (key[j] > T) keep |= 1u << j;
  if (need == c_eq) {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (key[j] == T) keep |= 1u << j;
  } else {
    // rare: more ties than room -> take the `need` smallest ids among them
    uint32_t taken = 0u;
    for (int n = 0; n < need; ++n) {
      unsigned long long best = ~0ull;
      int bj = -1;
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (key[j] == T && !(taken & (1u << j))) {
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
      if (bj >= 0 && best == wbest && (int)(__ffs(owner) - 1) == lane) { taken |= 1u << bj; keep |= 1u << bj; }
    }
  }
  int base = 0;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const bool kp = (keep >> j) & 1u;
    const unsigned b = __ballot_sync(0xffffffffu, kp);
    if (kp) list[base + __popc(b & ((1u << lane) - 1u))] = make_uint2(__float_as_uint(key2f(key[j])), pos[j]);
    base += __popc(b);
  }
  __syncwarp();
  *kth_key = T;
  return P.k;
}

template <int EPL>
__global__ void __launch_bounds__(kThreads, 1)
flat_ip_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_p, FipParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment for SWIZZLE_128B tiles, derived by pointer arithmetic on the __shared__ array so the
  // compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  FipShared* S = reinterpret_cast<FipShared*>(smem + (size_t)kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = P.dim / 64;
  const int n_items = P.n_qblocks * P.n_ranges;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_p);
    for (int s = 0; s < kStages; ++s) { mbar_init(&S->full[s], 1); mbar_init(&S->empty[s], 1); }
    for (int s = 0; s < kAccSlots; ++s) { mbar_init(&S->accfull[s], 1); mbar_init(&S->accempty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&S->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = S->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rg = item / P.n_qblocks, qb = item % P.n_qblocks;  // range-major: co-running CTAs share passages
        const int t0 = rg * P.tiles_per_range, t1 = min(P.n_tiles, t0 + P.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&S->empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&S->full[stage], (uint32_t)kStageBytes);
            uint8_t* st = smem + (size_t)stage * kStageBytes;
            tma_load_2d(&tmap_q, st, &S->full[stage], kb * 64, qb * BM, kEvictLast);
            tma_load_2d(&tmap_p, st + kABytes, &S->full[stage], kb * 64, t * BN, kEvictFirst);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc((uint32_t)P.fmt, BM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, accphase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rg = item / P.n_qblocks;
        const int t0 = rg * P.tiles_per_range, t1 = min(P.n_tiles, t0 + P.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&S->accempty[acc], accphase ^ 1u);
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&S->full[stage], phase);
            tc_fence_after_sync();
            const uint32_t a = smem_u32(smem + (size_t)stage * kStageBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_d, make_sw128_kmajor_desc(a + k * 32), make_sw128_kmajor_desc(a + kABytes + k * 32), idesc,
                       (uint32_t)((kb | k) != 0));
            umma_commit(&S->empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&S->accfull[acc]);
          if (++acc == kAccSlots) { acc = 0; accphase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------- epilogue: filter + top-k lists ---------------------------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // query row inside the block == TMEM lane
    int acc = 0;
    uint32_t accphase = 0;
    uint2* my_list = P.lists + ((size_t)blockIdx.x * BM + row) * P.cap;
    uint2* warp_lists = P.lists + ((size_t)blockIdx.x * BM + quarter * 32) * P.cap;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rg = item / P.n_qblocks, qb = item % P.n_qblocks;  // range-major: co-running CTAs share passages
      const int t0 = rg * P.tiles_per_range, t1 = min(P.n_tiles, t0 + P.tiles_per_range);
      const int64_t q = (int64_t)qb * BM + row;
      const bool live = q < P.nq;
      int cnt = 0;
      uint32_t tau_key = live ? P.tau_glob[q] : 0xffffffffu;  // dead rows accept nothing
      for (int t = t0; t < t1; ++t) {
        if (live) tau_key = max(tau_key, P.tau_glob[q]);
        float tau = key2f(tau_key);
        mbar_wait(&S->accfull[acc], accphase);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
        const int64_t col0 = (int64_t)t * BN;
        const bool ragged = col0 + BN > P.n_pass;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == BN / 32 - 1) {  // accumulator fully read: hand the slot back before any slow path
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->accempty[acc]);
          }
          // room for 32 appends per row is guaranteed by compacting whenever cnt > cap - 32
          const unsigned full_rows = __ballot_sync(0xffffffffu, cnt > P.cap - 32);
          for (unsigned m = full_rows; m; m &= m - 1) {
            const int rr = __ffs(m) - 1;
            const int c_rr = __shfl_sync(0xffffffffu, cnt, rr);
            uint32_t kth;
            const int nc = compact_row<EPL>(P, warp_lists + (size_t)rr * P.cap, c_rr, lane, &kth);
            if (lane == rr) {
              cnt = nc;
              tau_key = max(tau_key, kth);
              tau = key2f(tau_key);
              atomicMax(P.tau_glob + q, tau_key);
            }
          }
          const uint32_t pbase = (uint32_t)(col0 + c * 32);
          // steady state: almost no (row, 32-column group) holds a candidate -> one FMNMX3 chain and a skip
          float gmax = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), __uint_as_float(r[2]));
#pragma unroll
          for (int j = 3; j + 1 < 32; j += 2) gmax = fmaxf(fmaxf(gmax, __uint_as_float(r[j])), __uint_as_float(r[j + 1]));
          gmax = fmaxf(gmax, __uint_as_float(r[31]));
          if (gmax >= tau) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(r[j]);
              bool pass = s >= tau;
              if (ragged) pass = pass && (int64_t)(pbase + j) < P.n_pass;
              if (pass) { my_list[cnt] = make_uint2(r[j], pbase + j); ++cnt; }
            }
          }
        }
        if (++acc == kAccSlots) { acc = 0; accphase ^= 1u; }
      }
      // item done: final compaction of every row, then publish (score, id) candidates for the merge
      __syncwarp();
      for (int rr = 0; rr < 32; ++rr) {
        const int c_rr = __shfl_sync(0xffffffffu, cnt, rr);
        uint32_t kth;
        const int nc = compact_row<EPL>(P, warp_lists + (size_t)rr * P.cap, c_rr, lane, &kth);
        const int64_t qq = (int64_t)qb * BM + quarter * 32 + rr;
        if (lane == rr) {
          cnt = nc;
          if (live && nc == P.k) atomicMax(P.tau_glob + q, kth);
        }
        if (qq < P.nq) {
          const uint2* lst = warp_lists + (size_t)rr * P.cap;
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
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

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

__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ cand_scores,
                                                         const int64_t* __restrict__ cand_ids, int64_t nq, int L,
                                                         int Lpow2, int k, float* __restrict__ out_scores,
                                                         int64_t* __restrict__ out_ids) {
  extern __shared__ __align__(16) uint8_t msm[];
  Cand* c = reinterpret_cast<Cand*>(msm);
  for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
    __syncthreads();
    for (int e = threadIdx.x; e < Lpow2; e += blockDim.x) {
      Cand v;
      if (e < L) {
        v.s = cand_scores[q * L + e];
        v.id = cand_ids[q * L + e];
        v.valid = (v.id >= 0 && v.s == v.s && v.s != -INFINITY) ? 1 : 0;
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
      out_scores[q * k + e] = ok ? c[e].s : -3.4028234663852886e38f;  // faiss's "no result" convention
      out_ids[q * k + e] = ok ? c[e].id : -1;
    }
  }
}

struct Plan {
  int n_qblocks, n_tiles, n_ranges, tiles_per_range, grid, kpad, cap, epl;
};

Plan make_plan(int64_t nq, int64_t n_pass, int k, int sm_count) {
  Plan pl;
  pl.n_qblocks = (int)((nq + BM - 1) / BM);
  pl.n_tiles = (int)((n_pass + BN - 1) / BN);
  pl.kpad = (k + 31) / 32 * 32;
  pl.epl = k <= 128 ? 16 : 32;
  pl.cap = 32 * pl.epl;
  // number of passage ranges: every range restarts its threshold at -inf and pays ~log(range/k) list
  // compactions per query, so take the SMALLEST count whose item grid fills the SMs to >= 88 % (or the best
  // fill available).  MMB200_FLATIP_RANGES overrides it for experiments.
  int best_r = 1;
  double best_eff = -1.0;
  const int max_r = std::max(1, std::min(kMaxRanges, pl.n_tiles));
  double effs[kMaxRanges + 1];
  for (int r = 1; r <= max_r; ++r) {
    const int64_t items = (int64_t)pl.n_qblocks * r;
    const int64_t g = std::min<int64_t>(sm_count, items);
    const int64_t waves = (items + g - 1) / g;
    effs[r] = (double)items / (double)(waves * sm_count);
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
  pl.grid = (int)std::min<int64_t>(sm_count, (int64_t)pl.n_qblocks * pl.n_ranges);
  return pl;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t workspace_bytes(const Plan& pl, int64_t nq) {
  return align256((size_t)pl.grid * BM * pl.cap * sizeof(uint2)) + align256((size_t)nq * sizeof(uint32_t)) +
         align256((size_t)nq * pl.n_ranges * pl.kpad * sizeof(float)) +
         align256((size_t)nq * pl.n_ranges * pl.kpad * sizeof(int64_t));
}

__global__ void fill_u32(uint32_t* p, int64_t n, uint32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

int launch_merge(const float* cand_scores, const int64_t* cand_ids, int64_t nq, int L, int k, float* out_scores,
                 int64_t* out_ids, const DeviceInfo& dev, cudaStream_t stream) {
  int lp = 1;
  while (lp < L) lp <<= 1;
  lp = std::max(lp, 2);
  const size_t smem = (size_t)lp * sizeof(Cand);
  if (smem > (size_t)dev.max_smem_optin) {
    set_error("topk merge: too many candidates per query for one shared-memory sort (" + std::to_string(L) + ")");
    return MMB200_ERR_UNSUPPORTED;
  }
  MMB_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(nq, (int64_t)dev.sm_count * 4);
  topk_merge_kernel<<<grid, 256, smem, stream>>>(cand_scores, cand_ids, nq, L, lp, k, out_scores, out_ids);
  MMB_CHECK_CUDA(cudaGetLastError());
  return MMB200_OK;
}

}  // namespace

}  // namespace mmb

extern "C" int64_t mmb200_flat_ip_workspace_bytes(int64_t nq, int64_t n_pass, int32_t k) {
  using namespace mmb;
  if (nq <= 0 || n_pass <= 0 || k <= 0 || k > 256) return 0;
  DeviceInfo dev;
  if (current_device_info(&dev)) return -1;
  return (int64_t)workspace_bytes(make_plan(nq, n_pass, k, dev.sm_count), nq);
}

extern "C" int mmb200_flat_ip_topk(const void* queries, const void* passages, const int64_t* ids, float* out_scores,
                                   int64_t* out_ids, void* workspace, int64_t workspace_bytes_given, int64_t nq,
                                   int64_t n_pass, int32_t dim, int32_t k, int32_t dtype, int64_t id_base,
                                   void* stream_) {
  using namespace mmb;
  MMB_REQUIRE(queries && passages && out_scores && out_ids && workspace, "null pointer");
  MMB_REQUIRE(nq > 0 && n_pass > 0, "need at least one query and one passage");
  MMB_REQUIRE(k >= 1 && k <= 256, "fused top-k supports 1 <= k <= 256");
  MMB_REQUIRE(dtype == MMB200_F16 || dtype == MMB200_BF16, "passage storage must be fp16 or bf16 (faiss useFloat16)");
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
  MMB_REQUIRE((size_t)workspace_bytes_given >= workspace_bytes(pl, nq), "workspace too small (see mmb200_flat_ip_workspace_bytes)");

  FipParams P{};
  uint8_t* w = static_cast<uint8_t*>(workspace);
  P.lists = reinterpret_cast<uint2*>(w);
  w += align256((size_t)pl.grid * BM * pl.cap * sizeof(uint2));
  P.tau_glob = reinterpret_cast<uint32_t*>(w);
  w += align256((size_t)nq * sizeof(uint32_t));
  P.cand_scores = reinterpret_cast<float*>(w);
  w += align256((size_t)nq * pl.n_ranges * pl.kpad * sizeof(float));
  P.cand_ids = reinterpret_cast<int64_t*>(w);
  P.ids = ids; P.id_base = id_base; P.nq = nq; P.n_pass = n_pass; P.dim = dim; P.k = k; P.kpad = pl.kpad; P.cap = pl.cap;
  P.n_qblocks = pl.n_qblocks; P.n_ranges = pl.n_ranges; P.tiles_per_range = pl.tiles_per_range; P.n_tiles = pl.n_tiles;
  P.fmt = dtype == MMB200_F16 ? kFmtF16 : kFmtBF16;

  fill_u32<<<64, 256, 0, stream>>>(P.tau_glob, nq, kKeyNegInf);
  MMB_CHECK_CUDA(cudaGetLastError());

  const CUtensorMapDataType tdt = dtype == MMB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tq, tp;
  {
    const uint64_t dims[2] = {(uint64_t)dim, (uint64_t)nq};
    const uint64_t strides[1] = {(uint64_t)dim * 2};
    const uint32_t box[2] = {64, BM};
    if (int rc = encode_tensor_map(&tq, tdt, 2, queries, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)dim, (uint64_t)n_pass};
    const uint64_t strides[1] = {(uint64_t)dim * 2};
    const uint32_t box[2] = {64, BN};
    if (int rc = encode_tensor_map(&tp, tdt, 2, passages, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  const size_t smem = (size_t)kStages * kStageBytes + sizeof(FipShared) + 1024;
  if (pl.epl == 16) {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(flat_ip_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flat_ip_tc_kernel<16><<<pl.grid, kThreads, smem, stream>>>(tq, tp, P);
  } else {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(flat_ip_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flat_ip_tc_kernel<32><<<pl.grid, kThreads, smem, stream>>>(tq, tp, P);
  }
  MMB_CHECK_CUDA(cudaGetLastError());
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
