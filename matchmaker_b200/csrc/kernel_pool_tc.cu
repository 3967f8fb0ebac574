// Cosine + RBF kernel pooling forward (KNRM / TK) on the tensor cores with fp32-grade accuracy -- FIRST generation,
// both MMA operands in shared memory.  Superseded by kernel_pool_ts.cu (document operand in tensor memory); kept behind
// MMB200_KP_VARIANT=ss as an in-library cross-check and for A/B timing (0.376 ms vs 0.283 ms on the TK shape).
//
// The contraction q_i . d_j needs ~1e-7 accuracy (the RBF exponent amplifies cosine error by up to
// (c-mu)/sigma^2) and there is no fp32 tcgen05.mma kind, so every operand is split
// x = hi + lo with hi = x & 0xffffe000 (exactly representable in TF32, whatever the hardware's rounding)
// and lo = x - hi (exact in fp32).  The two query halves are STACKED along the UMMA N dimension:
//
//     D[128 doc rows x 64] = Dhi[128 x K] * [Qhi; Qlo]^T  +  Dlo[128 x K] * [Qhi; Qlo]^T
//
// so column i holds (dhi+dlo).qhi_i and column 32+i holds (dhi+dlo).qlo_i: 2 MMAs per k-step instead of
// the usual 3 of "3xTF32", and the cosine is one register add in the epilogue.
//
// Per CTA (persistent, one per SM, 480 threads):
//   warp 0      TMA producer: fp32 K-chunks [128 doc rows x 32] (16 KB) + [32 query rows x 32] (4 KB),
//               SWIZZLE_128B, into a RAW ring (4 slots of 20 KB); rows >= Ld / columns >= D zero-filled by TMA
//   warps 2-6   convert (one thread per row): raw slot -> registers -> hi / lo written to a 2-slot OPERAND ring
//               (the raw slot is released as soon as the values are stored, not when the MMA retires), running sum of
//               squares for the L2 norms (the normalisation is applied to the accumulator, not the operands)
//   warp 1      tcgen05.mma kind::tf32 issuer, 4 accumulator slots of 64 TMEM columns
//   warps 7-14  epilogue.  Phase A: tcgen05.ld, add the two halves, scale by 1/(|q|+eps) 1/(|d|+eps), write
//               the cosine tile to shared memory TRANSPOSED-friendly (16-byte chunks XOR-swizzled) with
//               masked / padded document rows replaced by a sentinel that zeroes every kernel.
//               Phase B: lane = query row, warp = 16 document rows: K activations ex2(-((c-mu)a)^2)
//               accumulated in registers -- the sum over documents needs no shuffles.  One cross-warp
//               reduction per pair, then log / mask / Linear(K,1).
//
// Bound: HBM by bytes ((Lq+Ld)*D*4 per pair), co-limited by MUFU ex2 at K=21.  Falls back (handled=false)
// to the FFMA kernel for Lq > 32, K > 32, or when the cosine matrix itself is requested.
#include <algorithm>
#include <cstdlib>

#include "host_util.cuh"
#include "kernel_pool.cuh"
#include "masks.cuh"
#include "ptx.cuh"

namespace mmb {

namespace {

constexpr int kThreads = 480;
constexpr int kMaxRaw = 8;            // raw ring (TMA targets): 20 KB per slot
constexpr int kMaxOps = 4;            // operand ring (what the MMA reads): 24 KB per slot (40 KB with an explicit hi tile)
constexpr int kAcc = 4;
constexpr int kDxBytes = 128 * 128;   // [128 rows][32 fp32]
constexpr int kQxBytes = 32 * 128;    // [32 query rows][32 fp32]
constexpr int kRawBytes = kDxBytes + kQxBytes;          // 20 KB
constexpr int kQ64Bytes = 64 * 128;   // rows 0-31 Q hi, rows 32-63 Q lo
constexpr int kOpBytes = 2 * kDxBytes + kQ64Bytes;      // 40 KB: Dhi | Dlo | [Qhi;Qlo]
constexpr int kConvThreads = 160;
constexpr int kEpiThreads = 256;
constexpr float kSentinel = 1.0e4f;   // "cosine" of a masked row: every kernel underflows to exactly 0
constexpr float kTinyNorm = 1e-13f;
constexpr float kClampMin = 1e-10f;

struct KpShared {
  uint64_t raw_full[kMaxRaw];    // TMA -> convert
  uint64_t raw_empty[kMaxRaw];   // convert (160 arrivals) -> TMA
  uint64_t op_full[kMaxOps];     // convert (160 arrivals) -> MMA
  uint64_t op_empty[kMaxOps];    // tcgen05.commit -> convert
  uint64_t accfull[kAcc];
  uint64_t accempty[kAcc];
  uint32_t tmem_base;
  uint32_t pad;
  float rs_d[kAcc][128];
  float rs_q[kAcc][32];
  float mu[32], a[32], alpha[32], w[32];
  float pk[32];
  float qm[32];
  float lsm[32 * 32];
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int KB>
__global__ void __launch_bounds__(kThreads, 1)
kernel_pool_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d, KpParams P,
                      int n_raw, int raw_hi, int kOps) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment for SWIZZLE_128B tiles, derived by pointer arithmetic on the __shared__ array so the
  // compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ops = smem;                                                          // [kOps][Dhi | Dlo | Q64]
  // raw_hi mode: the MMA reads the document hi operand straight from the raw tile (the tensor core ignores the 13
  // low mantissa bits of a TF32 operand), so the operand slot only holds Dlo | [Qhi;Qlo] and the raw slot is
  // released by the MMA's commit instead of by the convert threads
  const int op_bytes = raw_hi ? kDxBytes + kQ64Bytes : kOpBytes;
  const int dlo_off = raw_hi ? 0 : kDxBytes;
  const int q64_off = dlo_off + kDxBytes;
  uint8_t* raws = smem + kOps * op_bytes;                                       // [n_raw][Dx | Qx]
  float* cs = reinterpret_cast<float*>(raws + (size_t)n_raw * kRawBytes);       // [2][128][32] cosine tiles
  // the end-of-pair scratch aliases the cosine tiles (free between the last phase B of a pair and the first
  // phase A of the next one; fenced by named barriers 5 and 4) so that the shared memory goes to raw slots
  float* spart = cs;                                                           // [8][KB][32]  (<= 32 KB)
  KpShared* S = reinterpret_cast<KpShared*>(cs + 2 * 128 * 32);
  float* lsm = S->lsm;                                                         // [KB][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (P.Ld + 127) / 128;
  const int nch = (P.D + 31) / 32;
  const int64_t per = P.B / gridDim.x, rem = P.B % gridDim.x;
  const int64_t p_begin = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int64_t p_end = p_begin + per + ((int64_t)blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_d);
    for (int s = 0; s < n_raw; ++s) { mbar_init(&S->raw_full[s], 1); mbar_init(&S->raw_empty[s], raw_hi ? 1 : kConvThreads); }
    for (int s = 0; s < kOps; ++s) { mbar_init(&S->op_full[s], kConvThreads); mbar_init(&S->op_empty[s], 1); }
    for (int s = 0; s < kAcc; ++s) { mbar_init(&S->accfull[s], 1); mbar_init(&S->accempty[s], 8); }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    const int t = threadIdx.x;
    const bool ok = t < P.K;
    S->mu[t] = ok ? P.mu[t] : 0.f;
    S->a[t] = ok ? sqrtf(0.5f * 1.4426950408889634f) / P.sigma[t] : 0.f;
    S->alpha[t] = ok ? (P.alpha ? P.alpha[t] : 1.f) : 1.f;
    S->w[t] = ok ? P.weight[t] : 0.f;
  }
  if (warp == 1) tmem_alloc(&S->tmem_base, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = S->tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t p = p_begin; p < p_end; ++p)
        for (int t = 0; t < tiles; ++t)
          for (int ck = 0; ck < nch; ++ck) {
            mbar_wait(&S->raw_empty[stage], phase ^ 1u);
            uint8_t* st = raws + (size_t)stage * kRawBytes;
            mbar_arrive_expect_tx(&S->raw_full[stage], (uint32_t)kRawBytes);
            tma_load_3d(&tmap_d, st, &S->raw_full[stage], ck * 32, t * 128, (int)p, kEvictFirst);
            tma_load_3d(&tmap_q, st + kDxBytes, &S->raw_full[stage], ck * 32, 0, (int)p, kEvictLast);
            if (++stage == n_raw) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kFmtTF32, 128, 64);
      int stage = 0, acc = 0, rslot = 0;
      uint32_t phase = 0, accphase = 0;
      for (int64_t p = p_begin; p < p_end; ++p)
        for (int t = 0; t < tiles; ++t) {
          mbar_wait(&S->accempty[acc], accphase ^ 1u);
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 64);
          for (int ck = 0; ck < nch; ++ck) {
            mbar_wait(&S->op_full[stage], phase);
            tc_fence_after_sync();
            const uint32_t base = smem_u32(ops + (size_t)stage * op_bytes);
            const uint32_t hi_base = raw_hi ? smem_u32(raws + (size_t)rslot * kRawBytes) : base;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 32 fp32 / UMMA_K(8)
              const uint64_t bq = make_sw128_kmajor_desc(base + q64_off + k * 32);
              umma_tf32(tmem_d, make_sw128_kmajor_desc(hi_base + k * 32), bq, idesc, (uint32_t)((ck | k) != 0));
              umma_tf32(tmem_d, make_sw128_kmajor_desc(base + dlo_off + k * 32), bq, idesc, 1u);
            }
            if (raw_hi) umma_commit(&S->raw_empty[rslot]);
            if (++rslot == n_raw) rslot = 0;
            umma_commit(&S->op_empty[stage]);
            if (++stage == kOps) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&S->accfull[acc]);
          if (++acc == kAcc) { acc = 0; accphase ^= 1u; }
        }
    }
  } else if (warp < 7) {
    // ------------------------------- convert: hi / lo split + norms --------------
    const int ct = threadIdx.x - 64;          // 0..159
    const bool is_q = ct >= 128;
    const int row = is_q ? ct - 128 : ct;     // row inside the tile
    const int sw = row & 7;
    int rs_ = 0, os_ = 0, acc = 0;      // raw slot, operand slot, accumulator slot
    uint32_t rphase = 0, ophase = 0;
    for (int64_t p = p_begin; p < p_end; ++p)
      for (int t = 0; t < tiles; ++t) {
        float4 ss4 = make_float4(0.f, 0.f, 0.f, 0.f);  // four partial sums: shorter rounding chains for |x|^2
        for (int ck = 0; ck < nch; ++ck) {
          mbar_wait(&S->raw_full[rs_], rphase);
          const uint8_t* raw = raws + (size_t)rs_ * kRawBytes;
          const uint8_t* xrow = (is_q ? raw + kDxBytes : raw) + row * 128;
          float4 x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4*>(xrow + ((c ^ sw) << 4));
#pragma unroll
          for (int c = 0; c < 8; ++c) {  // consumes every loaded value: the loads have landed once this has executed
            const float4 v = x[c];
            ss4.x = fmaf(v.x, v.x, ss4.x); ss4.y = fmaf(v.y, v.y, ss4.y); ss4.z = fmaf(v.z, v.z, ss4.z); ss4.w = fmaf(v.w, v.w, ss4.w);
          }
          const int raw_slot = rs_;
          if (++rs_ == n_raw) { rs_ = 0; rphase ^= 1u; }
          mbar_wait(&S->op_empty[os_], ophase ^ 1u);
          uint8_t* op = ops + (size_t)os_ * op_bytes;
          uint8_t* hrow = is_q ? op + q64_off + row * 128 : op + row * 128;
          uint8_t* lrow = is_q ? op + q64_off + (32 + row) * 128 : op + dlo_off + row * 128;
          const bool write_hi = is_q || !raw_hi;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int off = ((c ^ sw) << 4);
            const float4 v = x[c];
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
            if (write_hi) *reinterpret_cast<float4*>(hrow + off) = hi;
            *reinterpret_cast<float4*>(lrow + off) = lo;
          }
          if (ck == nch - 1) {
            const float rs = 1.0f / (sqrtf((ss4.x + ss4.y) + (ss4.z + ss4.w)) + kTinyNorm);
            if (is_q) S->rs_q[acc][row] = rs; else S->rs_d[acc][row] = rs;
          }
          // Release the raw slot only now: the stores above consumed every loaded value, so the loads have LANDED
          // (an arrive placed right after the LDS instructions is hoisted above their completion by ptxas -- the
          // TMA then overwrites rows that are still being read: observed as ~1 % corrupted pairs).
          if (!raw_hi) mbar_arrive(&S->raw_empty[raw_slot]);
          fence_proxy_async_smem();
          mbar_arrive(&S->op_full[os_]);
          if (++os_ == kOps) { os_ = 0; ophase ^= 1u; }
        }
        if (++acc == kAcc) acc = 0;
      }
  } else {
    // ------------------------------- epilogue ------------------------------------
    const int ew = warp - 7;            // 0..7
    const int qd = warp & 3;            // TMEM lane quarter
    const int h = ew >> 2;              // which 16 query columns of the 32 this warp extracts in phase A
    const int et = threadIdx.x - 7 * 32;  // 0..255
    const int dmt = P.d_mask ? P.mask_dtype : MMB200_MASK_NONE;
    const int qmt = P.q_mask ? P.mask_dtype : MMB200_MASK_NONE;
    int acc_slot = 0;
    uint32_t accphase = 0;
    int64_t tile_seq = 0;
    for (int64_t p = p_begin; p < p_end; ++p) {
      float acc[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) acc[k] = 0.f;
      uint64_t qraw = 0;
      if (ew == 0 && lane < P.Lq) qraw = qmt != MMB200_MASK_NONE ? mask_raw(P.q_mask, qmt, p * (int64_t)P.Lq + lane) : 1;
      for (int t = 0; t < tiles; ++t, ++tile_seq) {
        const int row = qd * 32 + lane;          // document row inside the tile
        const int g = t * 128 + row;
        uint64_t draw = 0;
        if (g < P.Ld) draw = dmt != MMB200_MASK_NONE ? mask_raw(P.d_mask, dmt, p * (int64_t)P.Ld + g) : 1;
        float* cbuf = cs + (tile_seq & 1) * (128 * 32);
        mbar_wait(&S->accfull[acc_slot], accphase);
        tc_fence_after_sync();
        {  // phase A
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc_slot * 64);
          uint32_t rh[16], rl[16];
          tmem_ld_32x32b_x16(taddr + 16 * h, rh);
          tmem_ld_32x32b_x16(taddr + 32 + 16 * h, rl);
          tmem_ld_wait();
          const bool valid = g < P.Ld && mask_test(draw, dmt);
          const float rsd = S->rs_d[acc_slot][row];
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float c = (__uint_as_float(rh[j]) + __uint_as_float(rl[j])) * rsd * S->rs_q[acc_slot][16 * h + j];
            v[j] = valid ? c : kSentinel;
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->accempty[acc_slot]);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int phys = (4 * h + cc) ^ (row & 7);
            *reinterpret_cast<float4*>(cbuf + row * 32 + phys * 4) = make_float4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
          }
        }
        if (++acc_slot == kAcc) { acc_slot = 0; accphase ^= 1u; }
        named_bar_sync(1, kEpiThreads);
        {  // phase B: lane = query row, this warp's 16 document rows
#pragma unroll 2
          for (int rr = 0; rr < 16; ++rr) {
            const int r = ew * 16 + rr;
            const float c = cbuf[r * 32 + (((lane >> 2) ^ (r & 7)) << 2) + (lane & 3)];
            if (c < 1.0e3f) {  // uniform across the warp: whole rows are masked
#pragma unroll
              for (int k = 0; k < KB; ++k) {
                const float u = (c - S->mu[k]) * S->a[k];
                acc[k] += ex2f(-u * u);
              }
            }
          }
        }
      }
      // ---- end of pair: S_ik = sum over the 8 warps, log, mask, per-kernel sums, score ----
      named_bar_sync(5, kEpiThreads);  // every warp is done reading the cosine tiles that spart/lsm alias
#pragma unroll
      for (int k = 0; k < KB; ++k) spart[(ew * KB + k) * 32 + lane] = acc[k];
      if (ew == 0) S->qm[lane] = (lane < P.Lq && mask_test(qraw, qmt)) ? 1.f : 0.f;
      named_bar_sync(2, kEpiThreads);
      for (int e = et; e < KB * 32; e += kEpiThreads) {
        const int k = e >> 5, i = e & 31;
        float Ssum = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) Ssum += spart[(w8 * KB + k) * 32 + i];
        float L = 0.f;
        if (k < P.K && i < P.Lq) {
          if (P.per_kernel_query) P.per_kernel_query[(p * P.Lq + i) * (int64_t)P.K + k] = Ssum;
          if (S->qm[i] != 0.f) L = P.log_scale * logf(fmaxf(Ssum * S->alpha[k], kClampMin));
        }
        lsm[k * 32 + i] = L;
      }
      named_bar_sync(3, kEpiThreads);
      for (int k = ew; k < KB; k += 8) {
        float v = lsm[k * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) S->pk[k] = v;
      }
      named_bar_sync(4, kEpiThreads);
      if (et < P.K && P.per_kernel) P.per_kernel[p * P.K + et] = S->pk[et];
      if (et == 0) {
        float s = 0.f;
        for (int k = 0; k < P.K; ++k) s = fmaf(S->pk[k], S->w[k], s);
        P.score[p] = s;
      }
      // spart / lsm / pk / qm are rewritten only after the next pair's tiles, i.e. after further barriers
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int KB>
int launch(const KpParams& P, const DeviceInfo& dev, cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& td) {
  static_assert(8 * KB * 32 <= 2 * 128 * 32, "end-of-pair scratch must fit inside the cosine tiles");
  const size_t fixed = (size_t)(2 * 128 * 32) * sizeof(float) + sizeof(KpShared) + 1024;
  // hardware fact (measured, profiles/r01_bringup_kernel_pool_tc.log): tcgen05 kind::tf32 ignores the 13 low
  // mantissa bits of its fp32 inputs, so the raw tile IS the hi operand.  MMB200_KP_RAW_HI=0 restores the explicit
  // hi tile (identical results).
  const char* env = getenv("MMB200_KP_RAW_HI");
  const int raw_hi = (env && env[0] == '0') ? 0 : 1;
  const size_t op_bytes = raw_hi ? (size_t)kDxBytes + kQ64Bytes : (size_t)kOpBytes;
  int kOps = 3;
  if (const char* e2 = getenv("MMB200_KP_OPS")) kOps = std::max(2, std::min(kMaxOps, atoi(e2)));
  const size_t avail = (size_t)dev.max_smem_optin - fixed - (size_t)kOps * op_bytes;
  int n_raw = std::min<int>(kMaxRaw, (int)(avail / kRawBytes));
  if (const char* e3 = getenv("MMB200_KP_RAW")) n_raw = std::max(2, std::min(n_raw, atoi(e3)));
  const size_t smem = (size_t)kOps * op_bytes + (size_t)n_raw * kRawBytes + fixed;
  if (n_raw < 2 || smem > (size_t)dev.max_smem_optin) {
    set_error("kernel_pool tcgen05: shared-memory plan does not fit");
    return MMB200_ERR_UNSUPPORTED;
  }
  MMB_CHECK_CUDA(cudaFuncSetAttribute(kernel_pool_tc_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(dev.sm_count, P.B);
  kernel_pool_tc_kernel<KB><<<grid, kThreads, smem, stream>>>(tq, td, P, n_raw, raw_hi, kOps);
  MMB_CHECK_CUDA(cudaGetLastError());
  return MMB200_OK;
}

}  // namespace

int kernel_pool_fwd_tc(const KpParams& P, const DeviceInfo& dev, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (P.Lq > 32 || P.K > 32 || P.cosine != nullptr || P.D % 4 != 0) return MMB200_OK;
  CUtensorMap tq, td;
  {
    const uint64_t dims[3] = {(uint64_t)P.D, (uint64_t)P.Lq, (uint64_t)P.B};
    const uint64_t strides[2] = {(uint64_t)P.D * 4, (uint64_t)P.Lq * P.D * 4};
    const uint32_t box[3] = {32, 32, 1};
    if (int rc = encode_tensor_map(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.q, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
      return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)P.D, (uint64_t)P.Ld, (uint64_t)P.B};
    const uint64_t strides[2] = {(uint64_t)P.D * 4, (uint64_t)P.Ld * P.D * 4};
    const uint32_t box[3] = {32, 128, 1};
    if (int rc = encode_tensor_map(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.d, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
  }
  *handled = true;
  if (P.K <= 12) return launch<12>(P, dev, stream, tq, td);
  if (P.K <= 24) return launch<24>(P, dev, stream, tq, td);
  return launch<32>(P, dev, stream, tq, td);
}

}  // namespace mmb
