// ColBERT max-sim, second-generation tcgen05 kernel ("queries on M"): the hot path for Lq <= 32.
//
// Why a second orientation: in maxsim.cu the accumulator is [128 document rows (TMEM lanes) x 32 query
// columns], so the max over document rows is a cross-lane reduction -- ~460 warp instructions per
// 128-row tile on warps that have nobody to hide latency behind (ncu: 28 % issue-slot use, epilogue-bound
// at 56 % of HBM peak).  Here the accumulator is transposed:
//
//     D[128 x TN] = Qrep[128 x dim] * Doc[TN x dim]^T        (one tcgen05.mma chain per document)
//
// TMEM lane = query token, TMEM column = document row, so the max over a document is a per-thread
// FMNMX chain over registers (no shuffles, no cross-warp combine, no barrier).  Rows 0..31 of Qrep are
// the query, rows 32..63 a second copy (so two epilogue warps, TMEM lane quarters 0 and 1, can each
// take every other pair); rows 64..127 read whatever follows in shared memory and are never looked at.
//
// The document mask is applied BY THE TENSOR CORE: one extra UMMA K-step multiplies a column of ones
// (query side) with a per-row penalty (document side): 0 for real tokens, -inf for padding and for the
// tile's rows past Ld, so masked rows can never win the max.  The reference's -1000 fill
// (matchmaker/models/colbert.py:69) only matters when it IS the max; that is reproduced exactly by one
// "virtual" document row (index Ld, zero data from TMA out-of-bounds fill) whose penalty is -1000 when
// the document has at least one masked position and -inf otherwise.  A helper warp writes the 6 KB
// penalty tile per document while TMA streams the 48 KB of token vectors.
//
// Per document: 1 TMA (box 64 x TN x KB), 4*KB+1 MMAs (M=128, N=TN<=256, K=16), TN/32 tcgen05.ld + TN/3
// FMNMX3 in one warp, one fp32 store.  HBM-bound by design.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "host_util.cuh"
#include "masks.cuh"
#include "maxsim.cuh"
#include "ptx.cuh"

namespace mmb {

namespace {

constexpr int kThreads = 192;  // warp 0,1: epilogue (TMEM lane quarters 0,1); 2: TMA; 3: MMA; 4: penalty writer; 5: spare
constexpr int kMaxStages = 4;
constexpr int kMaxAcc = 4;
constexpr int kQSlots = 2;
constexpr int kQRep = 2;                     // copies of the 32 query rows
constexpr int kQRows = 32;
constexpr int kQBlockBytes = kQRep * kQRows * 128;  // one k-block of the replicated query tile (8 KB)

struct QmShared {
  uint64_t full[kMaxStages];   // 2 arrivals: TMA producer (with tx bytes) + penalty writer
  uint64_t empty[kMaxStages];  // tcgen05.commit
  uint64_t qfull[kQSlots];
  uint64_t qempty[kQSlots];
  uint64_t accfull[kMaxAcc];
  uint64_t accempty[kMaxAcc];
  uint32_t tmem_base;
  uint32_t pad;
};

struct QmLaunch {
  int32_t kblocks;      // dim / 64 (1 or 2)
  int32_t tn;           // document rows per tile (multiple of 16, <= 256)
  int32_t tiles;        // tiles per document; tiles * tn >= Ld + 1
  int32_t stages;
  int32_t acc_slots;
  int32_t tmem_cols;
  int32_t fmt;
  int32_t doc_bytes;    // kblocks * tn * 128
  int32_t stage_bytes;  // doc_bytes + tn * 32
  uint16_t neg_inf, neg_1000, one;  // bit patterns in the storage dtype
};

// K-major operand with NO swizzle, 16 elements (32 B) along K: core matrices of 8 rows x 16 B;
// second K chunk at +128 B (LBO), next 8-row group at +256 B (SBO).
__device__ __forceinline__ uint64_t make_noswz_k16_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ uint32_t penalty_offset(int row) { return (uint32_t)((row >> 3) * 256 + (row & 7) * 16); }

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ int64_t pair_dmask_row_of(const MaxsimParams& P, int64_t p) {
  if (P.pair_dmask) return (int64_t)P.pair_dmask[p];
  return P.pair_d ? (int64_t)P.pair_d[p] : p;
}

// kArgmax: the training instantiation also tracks WHICH document row won each query token's max (what backward needs,
// matchmaker/models/colbert.py:71 through autograd): a compare + two selects per accumulator element instead of a third
// of an FMNMX3 -- ~9x the epilogue instructions, still a fraction of the ~2000 cycles a document's bytes take to arrive.
template <bool kArgmax>
__global__ void __launch_bounds__(kThreads, 1)
maxsim_qm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                 const __grid_constant__ CUtensorMap tmap_d16, MaxsimParams P, QmLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment for SWIZZLE_128B tiles, derived by pointer arithmetic on the __shared__ array so the
  // compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int qslot_bytes = L.kblocks * kQBlockBytes;
  uint8_t* q_base = smem;                                            // [kQSlots][kblocks][2 x 32 rows][128 B]
  uint8_t* stage_base = q_base + kQSlots * qslot_bytes;              // [stages][doc tile | penalty tile]
  uint8_t* ones_tile = stage_base + (size_t)L.stages * L.stage_bytes;  // [128 rows][16 elems], no swizzle
  QmShared* S = reinterpret_cast<QmShared*>(ones_tile + 4096);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int64_t per = P.n_pairs / gridDim.x, rem = P.n_pairs % gridDim.x;
  const int64_t p_begin = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int64_t p_end = p_begin + per + ((int64_t)blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_d);
    for (int s = 0; s < L.stages; ++s) { mbar_init(&S->full[s], 2); mbar_init(&S->empty[s], 1); }
    for (int s = 0; s < kQSlots; ++s) { mbar_init(&S->qfull[s], 1); mbar_init(&S->qempty[s], 1); }
    for (int s = 0; s < L.acc_slots; ++s) { mbar_init(&S->accfull[s], 1); mbar_init(&S->accempty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 4) {
    // zero every penalty tile (only element 0 of each row's first 16-B chunk is rewritten per document)
    // and build the ones tile: element (row, k=0) = 1, everything else 0
    for (int s = 0; s < L.stages; ++s) {
      uint4* pt = reinterpret_cast<uint4*>(stage_base + (size_t)s * L.stage_bytes + L.doc_bytes);
      for (int e = lane; e < L.tn * 2; e += 32) pt[e] = make_uint4(0, 0, 0, 0);
    }
    uint4* ot = reinterpret_cast<uint4*>(ones_tile);
    for (int e = lane; e < 256; e += 32) {
      // 16-B chunk e: row group e/16, k-chunk (e/8)%2, row e%8
      const bool first_chunk = ((e >> 3) & 1) == 0;
      ot[e] = make_uint4(first_chunk ? (uint32_t)L.one : 0u, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (P.rows_needed) {
    // ragged fetch leaves rows of a stage untouched: start from zeros so that stale rows are always finite
    // and the virtual row (index TN-1 >= Ld, only ever written by TMA zero fill) is zero
    for (int s = 0; s < L.stages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(stage_base + (size_t)s * L.stage_bytes);
      for (int e = threadIdx.x; e < L.doc_bytes / 16; e += kThreads) z[e] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == 3) tmem_alloc(&S->tmem_base, (uint32_t)L.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = S->tmem_base;

  if (warp == 2) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int64_t prev_q = -1;
      uint32_t qcount = 0;
      for (int64_t p = p_begin; p < p_end; ++p) {
        const int64_t qi = P.pair_q ? (int64_t)P.pair_q[p] : (p + P.pair_base) / P.docs_per_query;
        const int64_t di = P.pair_d ? (int64_t)P.pair_d[p] : p;
        if (qi != prev_q) {
          const uint32_t slot = qcount & 1u, use = qcount >> 1;
          mbar_wait(&S->qempty[slot], (use & 1u) ^ 1u);
          mbar_arrive_expect_tx(&S->qfull[slot], (uint32_t)(L.kblocks * kQRep * kQRows * 128));
          for (int kb = 0; kb < L.kblocks; ++kb)
            for (int r = 0; r < kQRep; ++r)
              tma_load_4d(&tmap_q, q_base + (size_t)slot * qslot_bytes + kb * kQBlockBytes + r * (kQRows * 128),
                          &S->qfull[slot], 0, 0, kb, (int)qi, kEvictLast);
          ++qcount;
          prev_q = qi;
        }
        const int need_rows = P.rows_needed ? P.rows_needed[di] : 0;
        for (int t = 0; t < L.tiles; ++t) {
          mbar_wait(&S->empty[stage], phase ^ 1u);
          uint8_t* dst = stage_base + (size_t)stage * L.stage_bytes;
          if (!P.rows_needed) {
            mbar_arrive_expect_tx(&S->full[stage], (uint32_t)L.doc_bytes);
            tma_load_4d(&tmap_d, dst, &S->full[stage], 0, t * L.tn, 0, (int)di, kEvictFirst);
          } else {
            // 16-row blocks up to the document's last unmasked row; the rest of the stage keeps stale
            // (finite) rows, which the penalty tile masks with -inf
            const int rows_here = min(max(need_rows - t * L.tn, 0), L.tn);
            const int nb = (rows_here + 15) >> 4;
            if (nb == 0) {
              mbar_arrive(&S->full[stage]);
            } else {
              mbar_arrive_expect_tx(&S->full[stage], (uint32_t)(nb * L.kblocks * 2048));
              for (int kb = 0; kb < L.kblocks; ++kb)
                for (int b16 = 0; b16 < nb; ++b16)
                  tma_load_4d(&tmap_d16, dst + kb * L.tn * 128 + b16 * 2048, &S->full[stage], 0, t * L.tn + b16 * 16, kb,
                              (int)di, kEvictFirst);
            }
          }
          if (++stage == L.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc((uint32_t)L.fmt, 128, (uint32_t)L.tn);
      const uint64_t ones_desc = make_noswz_k16_desc(smem_u32(ones_tile));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t accphase = 0;
      int64_t prev_q = -1;
      uint32_t qcount = 0;
      int cur_slot = 0;
      for (int64_t p = p_begin; p < p_end; ++p) {
        const int64_t qi = P.pair_q ? (int64_t)P.pair_q[p] : (p + P.pair_base) / P.docs_per_query;
        if (qi != prev_q) {
          if (prev_q >= 0) umma_commit(&S->qempty[cur_slot]);
          cur_slot = (int)(qcount & 1u);
          mbar_wait(&S->qfull[cur_slot], (qcount >> 1) & 1u);
          ++qcount;
          prev_q = qi;
        }
        const uint32_t qaddr = smem_u32(q_base + (size_t)cur_slot * qslot_bytes);
        for (int t = 0; t < L.tiles; ++t) {
          mbar_wait(&S->accempty[acc], accphase ^ 1u);
          mbar_wait(&S->full[stage], phase);
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * L.tn);
          const uint32_t daddr = smem_u32(stage_base + (size_t)stage * L.stage_bytes);
          for (int kb = 0; kb < L.kblocks; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_d, make_sw128_kmajor_desc(qaddr + kb * kQBlockBytes + k * 32),
                       make_sw128_kmajor_desc(daddr + kb * L.tn * 128 + k * 32), idesc, (uint32_t)((kb | k) != 0));
          }
          // + ones[128 x 16] * penalty[TN x 16]^T : adds penalty[row] to every query's score of that row
          umma_f16(tmem_d, ones_desc, make_noswz_k16_desc(daddr + L.doc_bytes), idesc, 1u);
          umma_commit(&S->empty[stage]);
          umma_commit(&S->accfull[acc]);
          if (++stage == L.stages) { stage = 0; phase ^= 1u; }
          if (++acc == L.acc_slots) { acc = 0; accphase ^= 1u; }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------- penalty writer -----------------------------
    const int dmt = P.d_mask ? P.mask_dtype : MMB200_MASK_NONE;
    int stage = 0;
    uint32_t phase = 0;
    uint64_t raw[8], raw_next[8];
    auto fetch = [&](int64_t p, int t, uint64_t (&dst)[8]) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = lane + 32 * k, g = t * L.tn + r;
        dst[k] = 1;
        if (dmt != MMB200_MASK_NONE && p < p_end && r < L.tn && g < P.Ld)
          dst[k] = mask_raw(P.d_mask, dmt, pair_dmask_row_of(P, p) * (int64_t)P.Ld + g);
      }
    };
    fetch(p_begin, 0, raw_next);
    for (int64_t p = p_begin; p < p_end; ++p) {
      bool any_masked = false;
      for (int t = 0; t < L.tiles; ++t) {
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = raw_next[k];
        {  // prefetch the mask words of the next tile
          int nt = t + 1;
          int64_t np = p;
          if (nt == L.tiles) { nt = 0; ++np; }
          fetch(np, nt, raw_next);
        }
        uint16_t pen[8];
        bool masked_here = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = lane + 32 * k, g = t * L.tn + r;
          const bool in_doc = r < L.tn && g < P.Ld;
          const bool ok = in_doc && mask_test(raw[k], dmt);
          masked_here |= in_doc && !ok;
          pen[k] = ok ? (uint16_t)0 : L.neg_inf;
        }
        any_masked |= __any_sync(0xffffffffu, masked_here);
        mbar_wait(&S->empty[stage], phase ^ 1u);
        uint8_t* pt = stage_base + (size_t)stage * L.stage_bytes + L.doc_bytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = lane + 32 * k, g = t * L.tn + r;
          if (r < L.tn) {
            const uint16_t v = (g == L.tiles * L.tn - 1) ? (any_masked ? L.neg_1000 : L.neg_inf) : pen[k];
            *reinterpret_cast<uint16_t*>(pt + penalty_offset(r)) = v;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->full[stage]);
        if (++stage == L.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp < 2) {
    // ------------------------------- epilogue ------------------------------------
    const int qmt = P.q_mask ? P.mask_dtype : MMB200_MASK_NONE;
    const int n32 = L.tn >> 5, tail16 = (L.tn & 16) != 0;
    for (int64_t n = warp; p_begin + n < p_end; n += 2) {
      const int64_t p = p_begin + n;
      const int64_t qi = P.pair_q ? (int64_t)P.pair_q[p] : (p + P.pair_base) / P.docs_per_query;
      uint64_t qraw = 0;
      if (lane < P.Lq) qraw = (qmt != MMB200_MASK_NONE) ? mask_raw(P.q_mask, qmt, qi * (int64_t)P.Lq + lane) : 1;
      float m = -INFINITY;
      int am = -1;   // row of the running maximum (first one on ties); stays -1 when nothing beats -inf
      for (int t = 0; t < L.tiles; ++t) {
        const int64_t u = n * L.tiles + t;  // tile sequence number inside this CTA
        const int acc = (int)(u % L.acc_slots);
        const uint32_t accphase = (uint32_t)((u / L.acc_slots) & 1);
        mbar_wait(&S->accfull[acc], accphase);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * L.tn);
        for (int c = 0; c < n32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          if constexpr (kArgmax) {
            const int col0 = t * L.tn + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]);
              const bool gt = v > m;
              m = gt ? v : m;
              am = gt ? col0 + j : am;
            }
            continue;
          }
          float a = max3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
          float b = max3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
#pragma unroll
          for (int j = 6; j + 3 < 32; j += 4) {
            a = max3(a, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
            b = max3(b, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          }
          m = max3(m, a, b);
          m = max3(m, __uint_as_float(r[30]), __uint_as_float(r[31]));
        }
        if (tail16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(taddr + n32 * 32, r);
          tmem_ld_wait();
          if constexpr (kArgmax) {
            const int col0 = t * L.tn + n32 * 32;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v = __uint_as_float(r[j]);
              const bool gt = v > m;
              m = gt ? v : m;
              am = gt ? col0 + j : am;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) m = max3(m, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->accempty[acc]);
      }
      if constexpr (kArgmax) {
        // rows >= Ld are the -inf padding and the virtual -1000 row: a max taken there carries no gradient (-1), like a
        // masked query token
        if (lane < P.Lq) P.argmax[p * (int64_t)P.Lq + lane] = (mask_test(qraw, qmt) && am < P.Ld) ? am : -1;
      }
      float total = mask_test(qraw, qmt) ? m : 0.f;  // lanes >= Lq carry qraw = 0
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) P.out[p] = total;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)L.tmem_cols);
  }
}

}  // namespace

// rows_needed[di] = 1 + last unmasked row (one warp per document)
__global__ void __launch_bounds__(256) rows_needed_kernel(const void* __restrict__ d_mask, int mask_dtype,
                                                          int32_t* __restrict__ rows_needed, int64_t n_d, int Ld) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_d) return;
  int last = 0;
  if (!d_mask) last = Ld;
  else
    for (int j = lane; j < Ld; j += 32)
      if (mask_at(d_mask, mask_dtype, w * Ld + j)) last = j + 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) rows_needed[w] = last;
}

int maxsim_rows_needed_launch(const void* d_mask, int mask_dtype, int32_t* rows_needed, int64_t n_d, int Ld,
                              cudaStream_t stream) {
  if (n_d == 0) return MMB200_OK;
  rows_needed_kernel<<<(unsigned)((n_d + 7) / 8), 256, 0, stream>>>(d_mask, d_mask ? mask_dtype : MMB200_MASK_NONE, rows_needed, n_d, Ld);
  MMB_CHECK_CUDA(cudaGetLastError());
  return MMB200_OK;
}

// Returns MMB200_OK with *handled = false when the shape is outside this kernel's envelope.
int maxsim_qm_launch(const MaxsimParams& P, int dtype, const DeviceInfo& dev, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (dtype != MMB200_F16 && dtype != MMB200_BF16) return MMB200_OK;
  if (P.Lq > kQRows || (P.dim != 64 && P.dim != 128)) return MMB200_OK;
  if (P.rows_needed && (P.pair_d || P.pair_dmask)) return MMB200_OK;  // rows_needed is indexed by the implicit doc id
  if ((reinterpret_cast<uintptr_t>(P.q) | reinterpret_cast<uintptr_t>(P.d)) & 15) return MMB200_OK;
  QmLaunch L;
  L.kblocks = P.dim / 64;
  const int rows = P.Ld + 1;  // + the virtual row that carries the reference's -1000 fill
  L.tiles = (rows + 255) / 256;
  L.tn = (((rows + L.tiles - 1) / L.tiles) + 15) / 16 * 16;
  L.doc_bytes = L.kblocks * L.tn * 128;
  L.stage_bytes = L.doc_bytes + L.tn * 32;
  L.acc_slots = std::min(kMaxAcc, 512 / L.tn);
  L.tmem_cols = 32;
  while (L.tmem_cols < L.acc_slots * L.tn) L.tmem_cols <<= 1;
  L.fmt = dtype == MMB200_F16 ? kFmtF16 : kFmtBF16;
  if (dtype == MMB200_F16) { L.neg_inf = 0xFC00; L.neg_1000 = 0xE3D0; L.one = 0x3C00; }
  else { L.neg_inf = 0xFF80; L.neg_1000 = 0xC47A; L.one = 0x3F80; }
  const int fixed = kQSlots * L.kblocks * kQBlockBytes + 4096 + (int)sizeof(QmShared) + 1024;
  L.stages = std::min(kMaxStages, (dev.max_smem_optin - fixed) / L.stage_bytes);
  if (L.stages < 2 || L.acc_slots < 2) return MMB200_OK;
  const size_t smem_bytes = (size_t)L.stages * L.stage_bytes + fixed;

  const CUtensorMapDataType tdt = dtype == MMB200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tq, td;
  {
    const uint64_t dims[4] = {64, (uint64_t)P.Lq, (uint64_t)L.kblocks, (uint64_t)P.n_q};
    const uint64_t strides[3] = {(uint64_t)P.dim * 2, 128, (uint64_t)P.Lq * P.dim * 2};
    const uint32_t box[4] = {64, (uint32_t)kQRows, 1, 1};
    if (int rc = encode_tensor_map(&tq, tdt, 4, P.q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
      return rc;
  }
  {
    const uint64_t dims[4] = {64, (uint64_t)P.Ld, (uint64_t)L.kblocks, (uint64_t)P.n_d};
    const uint64_t strides[3] = {(uint64_t)P.dim * 2, 128, (uint64_t)P.Ld * P.dim * 2};
    const uint32_t box[4] = {64, (uint32_t)L.tn, (uint32_t)L.kblocks, 1};
    if (int rc = encode_tensor_map(&td, tdt, 4, P.d, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  CUtensorMap td16;
  {
    const uint64_t dims[4] = {64, (uint64_t)P.Ld, (uint64_t)L.kblocks, (uint64_t)P.n_d};
    const uint64_t strides[3] = {(uint64_t)P.dim * 2, 128, (uint64_t)P.Ld * P.dim * 2};
    const uint32_t box[4] = {64, 16, 1, 1};
    if (int rc = encode_tensor_map(&td16, tdt, 4, P.d, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  *handled = true;
  const int grid = (int)std::min<int64_t>(dev.sm_count, P.n_pairs);
  if (P.argmax) {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(maxsim_qm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    maxsim_qm_kernel<true><<<grid, kThreads, smem_bytes, stream>>>(tq, td, td16, P, L);
  } else {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(maxsim_qm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    maxsim_qm_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(tq, td, td16, P, L);
  }
  MMB_CHECK_CUDA(cudaGetLastError());
  return MMB200_OK;
}

}  // namespace mmb
