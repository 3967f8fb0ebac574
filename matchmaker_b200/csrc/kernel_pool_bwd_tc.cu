// Cosine + RBF kernel pooling BACKWARD (KNRM / TK training step) on the tensor cores.
//
// Reference arithmetic: autograd through matchmaker/modules/cosine... -> models/knrm.py:52-84 /
// models/published/ecai20_tk.py:105-124 (restated in oracle/interaction_oracle.py, kernel_pool_backward).
//
// With c_ij the cosine, S_ik the pooled activations the forward saved and coef_ik = g w_k s / S_ik (0 for masked query
// terms and below the clamp):
//
//     G_ij  = dm_j * sum_k coef_ik K_ijk (mu_k - c_ij) / sigma_k^2                  (d loss / d c_ij)
//     dd^_j = sum_i G_ij q_i / (|q_i| + eps)          dq^_i = sum_j G_ij d_j / (|d_j| + eps)
//     dd_j  = dd^_j / (|d_j| + eps) - d_j (d^_j . dd^_j) / (|d_j| (|d_j| + eps)),  d^_j . dd^_j = sum_i G_ij c_ij
//     dq_i  likewise with q^_i . dq^_i = sum_j G_ij c_ij (per-warp partial sums, added in a fixed order).
//
// The two contractions are the whole cost of the FFMA backward (kernel_pool.cu: 1.13 ms for 1024 TK pairs, 0.08 of the
// HBM ceiling; this kernel: 0.173 ms, 0.50; 4096 pairs: 0.633 ms, 0.55).  Here both run as kind::tf32 UMMAs straight on the RAW fp32 tiles TMA delivers -- the 1 / norm factors
// are folded into G, so no converted copy of Q or D is ever written:
//
//   GEMM 1  dd^[128 doc rows x 64 features] = G1[128 x 32] (A from TENSOR MEMORY, thread = document row wrote it)
//                                             * Q box [32 query rows x 64 features] (B, MN-major: K runs over the rows;
//                                               32-bit MN-major operands exist only in the 32-byte-atom 128-byte swizzle,
//                                               so Q and D are fetched with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
//   GEMM 2  dq^[32(+96 idle) query rows x 64 features] += G2^T[32 x 128] (A, K-major shared memory, written by the same
//                                             threads) * D box [128 doc rows x 64 features] (B, MN-major)
//
// The cosines are not recomputed: the training forward (kernel_pool_ts_kernel<.., SAVE>) leaves them document-row-major
// together with the inverse norms (KpParams::saved), 26 KB per TK pair against 277 KB of embeddings.
// A stage of the ring = two document boxes + the two query boxes of the same 64 features (40 KB; the query boxes come
// again for every document tile: L2 hits) -- no resident query tile, four stages fit.  The document gradient of a stage is
// finished in place: the epilogue thread (= document row) combines the accumulator row with its own raw row of the box
// still in shared memory, overwrites it and the box leaves by TMA store; the query gradient walks its 32-feature chunks
// through four 4 KB staging boxes the same way at the end of the pair.
//
// Operand precision: G is rounded to tf32 (cvt.rna); the raw tiles are truncated by the tensor core (low 13 mantissa
// bits dropped, mean relative shrink 0.72 * 2^-11), which KpParams::tf32_comp undoes on average.  Gradients agree with
// fp64 autograd of the reference expression to 2-7e-4 of the largest entry (bar 1e-3, tests/test_kernel_pool_gpu.py);
// the reference itself trains under fp16 autocast (train.py:330-348).
//
// Per CTA (persistent, one per SM, 640 threads; registers re-dealt per warpgroup with setmaxnreg):
//   warp 0       TMA producer: stages of 2 document boxes [128 x 32] + 2 query boxes [32 x 32], four-slot ring
//   warp 1       UMMA issuer
//   warp 3       store agent: TMA stores of finished document-gradient stages, stage release
//   warp 2       per pair: coef table T_ik = coef_ik / sigma_k^2, d weight / d alpha partial sums (from S only)
//   warps 4-7    document-gradient epilogue (thread = document row = TMEM lane)
//   warps 8-15   G: warp (quarter, half) = 32 document rows x 16 query rows of every tile; tiles alternate between two
//                G slots (TMEM columns + shared-memory atoms) so the next tile is prepared while the current one streams
//   warp 16      query-gradient epilogue (lane = query row = TMEM lane): drains the accumulator stage by stage, own small
//                TMA pipeline (raw query box in, finished gradient box out)
//
// TMEM map (512 columns): [0,320) dq^ accumulator (D <= 320); [320,448) 2 dd^ accumulators of 64; [448,512) G1, 2 x 32.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "host_util.cuh"
#include "kernel_pool.cuh"
#include "masks.cuh"
#include "ptx.cuh"

namespace mmb {

namespace {

constexpr int kThreads = 640;
constexpr int kBoxBytes = 128 * 128;       // document box [128 rows][32 fp32], SWIZZLE_128B_ATOM_32B
constexpr int kStageBoxes = 2;
constexpr int kQBoxBytes = 32 * 128;       // query box [32 rows][32 fp32]
constexpr int kG2AtomBytes = 32 * 128;     // G2^T atom: 32 query rows x 32 document rows (K-major, 128-byte rows)
constexpr int kG2Bytes = 4 * kG2AtomBytes; // one 128-row tile
constexpr int kMaxStages = 8;
constexpr int kDqSlots = 4;               // staging boxes of the query-gradient epilogue (load, finish in place, store)
constexpr int kDdAcc = 2;
constexpr int kColDq = 0, kColDd = 320, kColG1 = 448;
constexpr int kMaxD = 320;
// setmaxnreg budgets per warpgroup; the launch allocates 640 x 96 registers, the re-deal uses (56 + 80 + 120 + 120 + 88) x 128
// = 59392 of those 61440.  (A second set of four document-epilogue warps -- 768 threads, 48/72/112/112/72/64 registers --
// halved the epilogue's share of a stage but measured 2 % slower overall: same-box A/B, profiles/r02_kernel_pool_bwd_investigation.md.)
constexpr int kRegsLight = 56, kRegsDd = 80, kRegsG = 120, kRegsDq = 88;
constexpr int kDdWarp0 = 4, kGWarp0 = 8, kDqWarp = 16;
constexpr int kDdWarps = 4;

template <int KBP>
struct BwShared {
  uint64_t raw_full[kMaxStages], raw_empty[kMaxStages];
  uint64_t store_ready[kMaxStages];   // document epilogue -> store agent: the stage holds finished gradient boxes
  uint64_t acc_full[kDdAcc], acc_empty[kDdAcc];
  uint64_t g_full[2], g_empty[2];
  uint64_t coef_full[2], coef_empty[2];
  uint64_t dq_full;
  uint64_t dqs_full[kDqSlots];   // TMA -> query epilogue: the query box of a chunk has landed in its staging slot
  uint64_t dq_empty[10];   // per stage of a tile: its 64 (32) accumulator columns have been drained by the query epilogue
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float T[2][KBP][32];   // [pair parity][kernel][query row]: coef_ik / sigma_k^2
  float rsq[2][32];                  // 1 / (|q_i| + eps) of the pair
  float rowsc[2][3][128];            // [G slot][rsd | projection, query rows 0-15 | 16-31][tile row]
  float ci_part[4][8][16];           // [pair & 3][G warp][query row of its half]: sum_j G_ij c_ij over the warp's rows
  float mu[32], a[32], is2[32], sig2[32], alpha[32], w[32];
  float hgate[2][2][128];            // [G slot][query-row half][tile row]: d loss / d gate of the document term (GATE)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MMB200_ENABLE_PROF builds (python -m matchmaker_b200.build --prof) + MMB200_KPB_PROF=1: one thread per role of CTA 0
// accumulates the cycles it spends in each wait / phase; printed by the launcher.  Compiled out of the product build.
#ifdef MMB200_ENABLE_PROF
#define KPB_T(slot, stmt)                    \
  do {                                       \
    const long long t0_ = clock64();         \
    stmt;                                    \
    pc[slot] += clock64() - t0_;             \
  } while (0)
// stage trace: absolute clock of 7 events for stages 16..31 of CTA 0 (prof[64 + 7 * (n - 16) + event])
#define KPB_TRACE(n, ev)                                                                        \
  do {                                                                                          \
    if (prof && blockIdx.x == 0 && (n) >= 16 && (n) < 32) prof[64 + 7 * ((n) - 16) + (ev)] = clock64(); \
  } while (0)
// tile trace: prof[176 + 8 * tseq + ev] for tiles 0..15 of CTA 0: 0 MMA before dq_empty/g_full waits, 1 MMA after them,
// 2 MMA tile issued, 3 G warp 8 before g_empty wait, 4 G start of compute, 5 G end of compute, 6 coef table of the pair ready
#define KPB_TTRACE(ts, ev)                                                                    \
  do {                                                                                        \
    if (prof && blockIdx.x == 0 && (ts) < 16) prof[176 + 8 * (ts) + (ev)] = clock64();       \
  } while (0)
#else
#define KPB_T(slot, stmt) stmt
#define KPB_TTRACE(ts, ev) \
  do {                     \
  } while (0)
#define KPB_TRACE(n, ev) \
  do {                   \
  } while (0)
#endif

template <int KB, bool GATE>
__global__ void __launch_bounds__(kThreads, 1)
kernel_pool_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                          const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dd,
                          KpParams P, int n_stages, int stage_boxes, int mn_sbo, long long* prof) {
#ifdef MMB200_ENABLE_PROF
  long long pc[6] = {0, 0, 0, 0, 0, 0};
  const long long t_start = clock64();
  unsigned long long t_ns;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns));
#endif
  constexpr int KBP = (KB + 3) & ~3;
  using Shared = BwShared<KBP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nch = (P.D + 31) / 32;
  const int stage_bytes = stage_boxes * (kBoxBytes + kQBoxBytes);   // [document boxes | the query boxes of the same features]
  const int stage_q_off = stage_boxes * kBoxBytes;
  // G2 first: the M = 128 UMMA reads 96 idle A rows past each 32-row atom (up to 12 KB past the slot) -- into the ring
  uint8_t* g2 = smem;                                         // [2][4 atoms][32 rows][128 B]
  uint8_t* ring = g2 + 2 * kG2Bytes;                          // [n_stages][2 boxes]
  uint8_t* dqs = ring + (size_t)n_stages * stage_bytes;       // [kDqSlots] query boxes of the query-gradient epilogue
  Shared* S = reinterpret_cast<Shared*>(dqs + kDqSlots * kQBoxBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (P.Ld + 127) / 128;
  const int spt = (nch + stage_boxes - 1) / stage_boxes;      // stages per tile
  const int64_t per = P.B / gridDim.x, rem = P.B % gridDim.x;
  const int64_t p_begin = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int64_t p_end = p_begin + per + ((int64_t)blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_q);
    prefetch_tensormap(&tmap_d);
    prefetch_tensormap(&tmap_dd);
    prefetch_tensormap(&tmap_dq);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&S->raw_full[s], 1);
      mbar_init(&S->raw_empty[s], 1);
      mbar_init(&S->store_ready[s], kDdWarps);
    }
    for (int s = 0; s < kDdAcc; ++s) { mbar_init(&S->acc_full[s], 1); mbar_init(&S->acc_empty[s], kDdWarps); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S->g_full[s], 8);
      mbar_init(&S->g_empty[s], 1 + kDdWarps);      // tcgen05.commit of the tile's last stage + the document-epilogue warps
      mbar_init(&S->coef_full[s], 1);
      mbar_init(&S->coef_empty[s], 8);
    }
    mbar_init(&S->dq_full, 1);
    for (int s = 0; s < kDqSlots; ++s) mbar_init(&S->dqs_full[s], 1);
    for (int s = 0; s < 10; ++s) mbar_init(&S->dq_empty[s], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    const int t = threadIdx.x;
    const bool ok = t < P.K;
    const float sg = ok ? P.sigma[t] : 1.f;
    S->mu[t] = ok ? P.mu[t] : 0.f;
    S->a[t] = ok ? sqrtf(0.5f * 1.4426950408889634f) / sg : 0.f;
    S->is2[t] = ok ? 1.0f / (sg * sg) : 0.f;
    S->sig2[t] = ok ? sg * sg : 0.f;
    S->alpha[t] = ok ? (P.alpha ? P.alpha[t] : 1.f) : 1.f;
    S->w[t] = ok ? P.weight[t] : 0.f;
  }
  if (warp == 1) tmem_alloc(&S->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = S->tmem_base;

  if (warp < 4) {
    setmaxnreg_dec<kRegsLight>();
    if (warp == 0) {
      // ------------------------------- TMA producer -------------------------------
      if (lane == 0) {
        int slot = 0;
        uint32_t ph = 0;
        for (int64_t p = p_begin; p < p_end; ++p) {
          for (int t = 0; t < tiles; ++t)
            for (int s = 0; s < spt; ++s) {
              const int nb = min(stage_boxes, nch - s * stage_boxes);
              KPB_T(1, mbar_wait<true>(&S->raw_empty[slot], ph ^ 1u));
              KPB_TRACE((int)(p - p_begin) * tiles * spt + t * spt + s, 0);
              mbar_arrive_expect_tx(&S->raw_full[slot], (uint32_t)(nb * (kBoxBytes + kQBoxBytes)));
              for (int b = 0; b < nb; ++b) {
                tma_load_3d(&tmap_d, ring + (size_t)slot * stage_bytes + (size_t)b * kBoxBytes, &S->raw_full[slot],
                            (s * stage_boxes + b) * 32, t * 128, (int)p, kEvictFirst);
                // the query boxes of the same features travel with the stage (a second time per document tile: L2 hits) --
                // no resident query tile, the shared memory it took is a fourth ring slot
                tma_load_3d(&tmap_q, ring + (size_t)slot * stage_bytes + stage_q_off + (size_t)b * kQBoxBytes, &S->raw_full[slot],
                            (s * stage_boxes + b) * 32, 0, (int)p, kEvictLast);
              }
              if (++slot == n_stages) { slot = 0; ph ^= 1u; }
            }
        }
      }
    } else if (warp == 1) {
      // ------------------------------- UMMA issuer --------------------------------
      int slot = 0, a = 0;
      uint32_t ph = 0, aph = 0;
      int tseq = 0;
      for (int64_t p = p_begin; p < p_end; ++p) {
        const int pi = (int)(p - p_begin);
        tc_fence_after_sync();
        for (int t = 0; t < tiles; ++t, ++tseq) {
          const int g = tseq & 1;
          if (lane == 0) KPB_TTRACE(tseq, 0);
          KPB_T(2, mbar_wait<true>(&S->g_full[g], (uint32_t)((tseq >> 1) & 1)));
          if (lane == 0) KPB_TTRACE(tseq, 1);
          tc_fence_after_sync();
          const int ksteps = (min(128, P.Ld - t * 128) + 7) >> 3;
          const uint32_t g1col = tmem_base + (uint32_t)(kColG1 + g * 32);
          const uint32_t g2addr = smem_u32(g2 + (size_t)g * kG2Bytes);
          for (int s = 0; s < spt; ++s) {
            const int nb = min(stage_boxes, nch - s * stage_boxes);
            const uint32_t idesc = make_idesc(kFmtTF32, 128, (uint32_t)(32 * nb)) | kIdescBMajorMN;
            // first tile of a pair overwrites the dq^ columns of this stage: the query epilogue of the previous pair drains
            // them stage by stage, so the UMMAs follow right behind it instead of waiting for the whole accumulator
            if (t == 0) KPB_T(1, mbar_wait<true>(&S->dq_empty[s], (uint32_t)((pi & 1) ^ 1)));
            KPB_T(3, mbar_wait<true>(&S->raw_full[slot], ph));
            if (lane == 0) KPB_TRACE(tseq * spt + s, 1);
            KPB_T(4, mbar_wait<true>(&S->acc_empty[a], aph ^ 1u));
            tc_fence_after_sync();
            const uint32_t dd_acc = tmem_base + (uint32_t)(kColDd + a * 64);
            const uint32_t dq_acc = tmem_base + (uint32_t)(kColDq + s * stage_boxes * 32);
            const uint32_t dbox = smem_u32(ring + (size_t)slot * stage_bytes);
            const uint32_t qbox = dbox + (uint32_t)stage_q_off;
            const bool last_s = s == spt - 1;
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)   // GEMM 1: K = the 32 query rows
                umma_tf32_ts(dd_acc, g1col + (uint32_t)(8 * k), make_sw128x32_mnmajor_desc(qbox + (uint32_t)(k * 1024), kQBoxBytes, (uint32_t)mn_sbo),
                             idesc, (uint32_t)(k != 0));
              for (int k = 0; k < ksteps; ++k)   // GEMM 2: K = the document rows of the tile
                umma_tf32(dq_acc, make_sw128_kmajor_desc(g2addr + (uint32_t)((k >> 2) * kG2AtomBytes + (k & 3) * 32)),
                          make_sw128x32_mnmajor_desc(dbox + (uint32_t)(k * 1024), kBoxBytes, (uint32_t)mn_sbo), idesc, (uint32_t)((t | k) != 0));
              umma_commit(&S->acc_full[a]);
              if (last_s) umma_commit(&S->g_empty[g]);
              if (last_s && t == tiles - 1) umma_commit(&S->dq_full);
            }
            __syncwarp();
            if (lane == 0) KPB_TRACE(tseq * spt + s, 2);
            if (lane == 0 && s == spt - 1) KPB_TTRACE(tseq, 2);
            if (++slot == n_stages) { slot = 0; ph ^= 1u; }
            if (++a == kDdAcc) { a = 0; aph ^= 1u; }
          }
        }
      }
    } else if (warp == 3) {
      // ------------------------------- store agent ----------------------------------
      // finished document-gradient boxes leave by TMA; the stage goes back to the producer as soon as the store has READ
      // it (waiting here instead of in the epilogue keeps the epilogue warps off the store's latency)
      if (lane == 0) {
        int slot = 0;
        uint32_t ph = 0;
        for (int64_t p = p_begin; p < p_end; ++p)
          for (int t = 0; t < tiles; ++t)
            for (int s = 0; s < spt; ++s) {
              const int nb = min(stage_boxes, nch - s * stage_boxes);
              KPB_T(0, mbar_wait<true>(&S->store_ready[slot], ph));
              KPB_TRACE((int)(p - p_begin) * tiles * spt + t * spt + s, 5);
              for (int b = 0; b < nb; ++b)
                tma_store_3d(&tmap_dd, ring + (size_t)slot * stage_bytes + (size_t)b * kBoxBytes, (s * stage_boxes + b) * 32,
                             t * 128, (int)p);
              bulk_commit_group();
              KPB_T(1, bulk_wait_group_read<0>());
              KPB_TRACE((int)(p - p_begin) * tiles * spt + t * spt + s, 6);
              mbar_arrive(&S->raw_empty[slot]);
              if (++slot == n_stages) { slot = 0; ph ^= 1u; }
            }
        bulk_wait_group<0>();
      }
    } else if (warp == 2) {
      // ------------------------------- coef table, d weight, d alpha ----------------
      const int qmt = P.q_mask ? P.mask_dtype : MMB200_MASK_NONE;
      for (int64_t p = p_begin; p < p_end; ++p) {
        const int pi = (int)(p - p_begin), pb = pi & 1;
        mbar_wait<true>(&S->coef_empty[pb], (uint32_t)(((pi >> 1) & 1) ^ 1));
        const float g = P.grad_score[p];
        const bool qlive = lane < P.Lq && mask_at(P.q_mask, qmt, p * (int64_t)P.Lq + lane);
        float Sr[KBP];   // the row's pooled activations, loaded back to back (one memory latency per pair, not one per kernel)
#pragma unroll
        for (int k = 0; k < KBP; ++k) Sr[k] = (k < P.K && qlive) ? P.S[(p * P.Lq + lane) * (int64_t)P.K + k] : 1.f;
#pragma unroll
        for (int k = 0; k < KBP; ++k) {
          float cf = 0.f, Lv = 0.f, da = 0.f;
          if (k < P.K && qlive) {
            const float Sv = Sr[k];
            const float aS = Sv * S->alpha[k];
            Lv = P.log_scale * logf(fmaxf(aS, P.clamp_min));
            if (aS >= P.clamp_min) {   // torch.clamp passes the gradient at equality
              cf = g * S->w[k] * P.log_scale / Sv;
              da = g * S->w[k] * P.log_scale / S->alpha[k];
            }
          }
          S->T[pb][k][lane] = cf * S->is2[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            Lv += __shfl_xor_sync(0xffffffffu, Lv, o);
            da += __shfl_xor_sync(0xffffffffu, da, o);
          }
          if (lane == 0 && k < P.K) {
            P.ws_weight[p * P.K + k] = g * Lv;
            P.ws_alpha[p * P.K + k] = da;
          }
        }
        S->rsq[pb][lane] = P.saved[kp_saved_rsq_off(P.B, p, P.Ld) + lane];
        __syncwarp();
        if (lane == 0) KPB_TTRACE(pi * tiles, 6);
        if (lane == 0) mbar_arrive(&S->coef_full[pb]);
      }
    }
  } else if (warp < kGWarp0) {
    // ------------------------------- document-gradient epilogue ----------------------
    setmaxnreg_dec<kRegsDd>();
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    int slot = 0, a = 0;
    uint32_t aph = 0;
    int tseq = 0;
    for (int64_t p = p_begin; p < p_end; ++p)
      for (int t = 0; t < tiles; ++t, ++tseq) {
        const int g = tseq & 1;
        for (int s = 0; s < spt; ++s) {
          const int nb = min(stage_boxes, nch - s * stage_boxes);
          KPB_T(0, mbar_wait<true>(&S->acc_full[a], aph));
          if (warp == kDdWarp0 && lane == 0) KPB_TRACE(tseq * spt + s, 3);
          tc_fence_after_sync();
          const float rsd = S->rowsc[g][0][row] * P.tf32_comp;
          const float pr = S->rowsc[g][1][row] + S->rowsc[g][2][row];
          if constexpr (GATE) {
            if (s == 0 && P.grad_gate && t * 128 + row < P.Ld)
              P.grad_gate[p * (int64_t)P.Ld + t * 128 + row] = S->hgate[g][0][row] + S->hgate[g][1][row];
          }
          uint8_t* stage = ring + (size_t)slot * stage_bytes;
          for (int b = 0; b < nb; ++b) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(kColDd + a * 64 + 32 * b), r);
            tmem_ld_wait();
            uint8_t* boxp = stage + (size_t)b * kBoxBytes;
            // rows r and r + 4 share their 32-byte units in this layout: rows with bit 2 set visit the two 16-byte halves
            // of each unit in the other order, so the 8 rows of a quarter warp touch 8 different 16-byte slots
            const bool flip = row & 4;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              float4* cell = reinterpret_cast<float4*>(boxp + sw128x32_offset(row, cc ^ (flip ? 1 : 0)));
              const float4 dv = *cell;
              float4 o;
              o.x = fmaf(rsd, __uint_as_float(flip ? r[4 * (cc ^ 1) + 0] : r[4 * cc + 0]), -dv.x * pr);
              o.y = fmaf(rsd, __uint_as_float(flip ? r[4 * (cc ^ 1) + 1] : r[4 * cc + 1]), -dv.y * pr);
              o.z = fmaf(rsd, __uint_as_float(flip ? r[4 * (cc ^ 1) + 2] : r[4 * cc + 2]), -dv.z * pr);
              o.w = fmaf(rsd, __uint_as_float(flip ? r[4 * (cc ^ 1) + 3] : r[4 * cc + 3]), -dv.w * pr);
              *cell = o;
            }
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&S->acc_empty[a]);
            if (s == spt - 1) mbar_arrive(&S->g_empty[g]);   // rowsc[g] is free for the slot's next tile
            mbar_arrive(&S->store_ready[slot]);
            if (warp == kDdWarp0) KPB_TRACE(tseq * spt + s, 4);
          }
          if (++slot == n_stages) slot = 0;
          if (++a == kDdAcc) { a = 0; aph ^= 1u; }
        }
      }
  } else if (warp < kDqWarp) {
    // ------------------------------- G: d loss / d cosine -------------------------------
    // All 8 warps work on every tile: warp (qd, hh) owns document rows 32 qd .. + 31 (its TMEM lane quarter, its G2^T
    // atom) and the query rows 16 hh .. + 15.  Tiles alternate between the two G slots, so the next tile's G is computed
    // while the UMMAs stream the current one.
    setmaxnreg_inc<kRegsG>();
    const int hh = (warp - kGWarp0) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int dmt = P.d_mask ? P.mask_dtype : MMB200_MASK_NONE;
    int tseq = 0;
    for (int64_t p = p_begin; p < p_end; ++p) {
      const int pi = (int)(p - p_begin), pb = pi & 1;
      KPB_T(0, mbar_wait<true>(&S->coef_full[pb], (uint32_t)((pi >> 1) & 1)));
      float ci_lane = 0.f;   // after each tile's butterfly: sum over this warp's document rows of G_ij c_ij, query row (lane >> 1) & 15
      for (int t = 0; t < tiles; ++t, ++tseq) {
        const int g = tseq & 1;
        uint8_t* g2slot = g2 + (size_t)g * kG2Bytes + (size_t)qd * kG2AtomBytes;   // this warp's 32 document rows = one atom
        const int rows_t = min(128, P.Ld - t * 128);
        const bool warp_live = qd * 32 < ((rows_t + 7) & ~7);   // some row of this warp is read by GEMM 2
        const int j = t * 128 + row;
        const bool inb = j < P.Ld;
        bool valid = false;
        float rsd = 0.f;
        const float* crow = P.saved + kp_saved_cos_off(p, P.Ld) + (int64_t)j * 32 + 16 * hh;
        if (inb) {
          valid = mask_at(P.d_mask, dmt, p * (int64_t)P.Ld + j);
          rsd = P.saved[kp_saved_rsd_off(P.B, p, P.Ld) + j];
        }
        const bool any_valid = __any_sync(0xffffffffu, valid);
        float c[16], G[16];
        if (valid && 16 * hh < P.Lq) {
#pragma unroll
          for (int x4 = 0; x4 < 4; ++x4) {
            const float4 v = *reinterpret_cast<const float4*>(crow + 4 * x4);
            c[4 * x4] = v.x; c[4 * x4 + 1] = v.y; c[4 * x4 + 2] = v.z; c[4 * x4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int x = 0; x < 16; ++x) c[x] = 0.f;
        }
        if (warp == kGWarp0 && lane == 0) KPB_TTRACE(tseq, 3);
        KPB_T(1, mbar_wait<true>(&S->g_empty[g], (uint32_t)(((tseq >> 1) & 1) ^ 1)));
        tc_fence_after_sync();
        if (warp == kGWarp0 && lane == 0) KPB_TTRACE(tseq, 4);
#ifdef MMB200_ENABLE_PROF
        const long long t_g0 = clock64();
#endif
        if (warp_live) {
#pragma unroll
          for (int x = 0; x < 16; ++x) G[x] = 0.f;
          float H[4] = {0.f, 0.f, 0.f, 0.f};   // GATE: sum_i sum_k coef_ik K_ijk = d loss / d gate_j (four partial chains)
          if (any_valid && 16 * hh < P.Lq) {   // warp-uniform
            // kernel by kernel, the 16 query rows side by side: 16 independent chains per step, centre and width of the
            // kernel as warp-uniform operands, the coef row of the kernel as four broadcast 16-byte loads
#pragma unroll 3
            for (int k = 0; k < KB; ++k) {
              const float mu_k = S->mu[k], a_k = S->a[k];
              const float sig2_k = GATE ? S->sig2[k] : 0.f;
              const float4* Tk = reinterpret_cast<const float4*>(&S->T[pb][k][16 * hh]);
#pragma unroll
              for (int x4 = 0; x4 < 4; ++x4) {
                const float4 T4 = Tk[x4];
                const float Tv[4] = {T4.x, T4.y, T4.z, T4.w};
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                  const int x = 4 * x4 + y;
                  const float diff = mu_k - c[x];
                  const float u = diff * a_k;
                  const float te = Tv[y] * ex2f(-u * u);
                  G[x] = fmaf(te, diff, G[x]);
                  if constexpr (GATE) H[y] = fmaf(te, sig2_k, H[y]);
                }
              }
            }
          }
          float cpr = 0.f;
          float gc[16];
          float gate_j = 1.f;
          if constexpr (GATE) {
            const float gv = valid ? P.gate[p * (int64_t)P.Ld + j] : 0.f;
            gate_j = fmaxf(gv, 0.f);   // the forward counts a negative gate as 0: relu'(gate) = 0 there
            S->hgate[g][hh][row] = (valid && gv >= 0.f) ? (H[0] + H[1]) + (H[2] + H[3]) : 0.f;
          }
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            G[x] = valid ? G[x] * gate_j : 0.f;
            gc[x] = G[x] * c[x];
            cpr += gc[x];
          }
          // G1 (A of GEMM 1, tensor memory): G_ij / (|q_i| + eps), this thread's TMEM lane, columns 16 hh .. + 15
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            uint32_t g1[8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) g1[ii] = f32_to_tf32_rna(G[8 * h8 + ii] * S->rsq[pb][16 * hh + 8 * h8 + ii]);
            tmem_st_32x32b_x8(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(kColG1 + g * 32 + 16 * hh + 8 * h8), g1);
          }
          // G2^T (A of GEMM 2, shared memory, K-major rows of 32 document rows): [query row][this warp's lane]
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int i = 16 * hh + x;
            *reinterpret_cast<uint32_t*>(g2slot + i * 128 + ((((lane >> 2) ^ (i & 7))) << 4) + ((lane & 3) << 2)) =
                f32_to_tf32_rna(G[x] * rsd);
          }
          if (hh == 0) S->rowsc[g][0][row] = rsd;
          S->rowsc[g][1 + hh][row] = rsd * cpr * (rsd < 1e12f ? rsd : 0.f);   // this half of (d^_j . dd^_j) / |d_j|; 0 for a zero row
          // q^_i . dq^_i = sum_j G_ij c_ij for the query epilogue: add this warp's 32 document rows with a transposing
          // butterfly -- after the steps lane l holds query row (l >> 1) & 15 of this warp's half
          if (any_valid) {
            float v8[8], v4[4], v2[2];
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
            for (int x = 0; x < 8; ++x) v8[x] = (b4 ? gc[8 + x] : gc[x]) + __shfl_xor_sync(0xffffffffu, b4 ? gc[x] : gc[8 + x], 16);
#pragma unroll
            for (int x = 0; x < 4; ++x) v4[x] = (b3 ? v8[4 + x] : v8[x]) + __shfl_xor_sync(0xffffffffu, b3 ? v8[x] : v8[4 + x], 8);
#pragma unroll
            for (int x = 0; x < 2; ++x) v2[x] = (b2 ? v4[2 + x] : v4[x]) + __shfl_xor_sync(0xffffffffu, b2 ? v4[x] : v4[2 + x], 4);
            float v1 = (b1 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? v2[0] : v2[1], 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            ci_lane += v1;
          }
          tmem_st_wait();
        }
#ifdef MMB200_ENABLE_PROF
        pc[2] += clock64() - t_g0;
        if (warp == kGWarp0 && lane == 0) KPB_TTRACE(tseq, 5);
#endif
        if (t == tiles - 1) {
          // one partial per warp and pair; the query epilogue adds the four quarter partials in a fixed order (no atomics:
          // deterministic).  Ring of 4 pairs: the G warps run at most two tiles ahead of the UMMAs (g_empty), the UMMAs of
          // pair p + 1 start after the query epilogue of pair p has read the partials -- pair p + 4 cannot be here before
          if ((lane & 1) == 0) S->ci_part[pi & 3][warp - kGWarp0][lane >> 1] = ci_lane;
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->g_full[g]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->coef_empty[pb]);
    }
  } else {
    // ------------------------------- query-gradient epilogue (warp 16) ---------------
    setmaxnreg_dec<kRegsDq>();
    if (warp == kDqWarp) {
      // The query gradient is 36 KB per pair against 240 KB of document gradient: no resident query tile (its shared memory
      // is the ring's fourth slot), the epilogue walks the 32-feature chunks through four 4 KB staging boxes -- TMA load of
      // the raw query box (L2: the UMMAs have just streamed it) two chunks ahead, finish in place, TMA store.
      int gc = 0;   // chunk sequence number of this CTA: staging slot gc % 4, parity (gc / 4) & 1
      auto stage_in = [&](int64_t pp, int c, int gcc) {   // lane 0: the slot's last store (chunk gcc - 4) has been read
        mbar_arrive_expect_tx(&S->dqs_full[gcc % kDqSlots], (uint32_t)kQBoxBytes);
        tma_load_3d(&tmap_q, dqs + (size_t)(gcc % kDqSlots) * kQBoxBytes, &S->dqs_full[gcc % kDqSlots], c * 32, 0, (int)pp, kEvictNormal);
      };
      for (int64_t p = p_begin; p < p_end; ++p) {
        const int pi = (int)(p - p_begin);
        const float rsq = P.saved[kp_saved_rsq_off(P.B, p, P.Ld) + lane];
        if (lane == 0) {   // first two chunks on their way while the pair's last UMMAs finish
          bulk_wait_group_read<0>();
          stage_in(p, 0, gc);
          if (nch > 1) stage_in(p, 1, gc + 1);
        }
        KPB_T(0, mbar_wait<true>(&S->dq_full, (uint32_t)(pi & 1)));
        tc_fence_after_sync();
        // q^_i . dq^_i: the four document-row quarters of this query row's half, fixed order
        const int w0 = (lane >> 4) * 4, x = lane & 15;
        const float(*cp)[16] = S->ci_part[pi & 3];
        const float cq = (cp[w0][x] + cp[w0 + 1][x]) + (cp[w0 + 2][x] + cp[w0 + 3][x]);
        const float s1 = rsq * P.tf32_comp;
        const float s2 = rsq * cq * (rsq < 1e12f ? rsq : 0.f);   // (q^_i . dq^_i) / |q_i|, times 1 / (|q_i| + eps)
        for (int c = 0; c < nch; ++c, ++gc) {
          if (lane == 0 && c + 2 < nch) {
            bulk_wait_group_read<1>();   // the store of chunk c - 2 (same slot as c + 2) has read its box; chunk c - 1's may be pending
            stage_in(p, c + 2, gc + 2);
          }
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t)(kColDq + 32 * c), r);
          KPB_T(1, mbar_wait<true>(&S->dqs_full[gc % kDqSlots], (uint32_t)((gc / kDqSlots) & 1)));
          tmem_ld_wait();
          uint8_t* box = dqs + (size_t)(gc % kDqSlots) * kQBoxBytes;
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            float4* cell = reinterpret_cast<float4*>(box + sw128x32_offset(lane, cc));
            const float4 qv = *cell;
            float4 o;
            o.x = fmaf(s1, __uint_as_float(r[4 * cc + 0]), -qv.x * s2);
            o.y = fmaf(s1, __uint_as_float(r[4 * cc + 1]), -qv.y * s2);
            o.z = fmaf(s1, __uint_as_float(r[4 * cc + 2]), -qv.z * s2);
            o.w = fmaf(s1, __uint_as_float(r[4 * cc + 3]), -qv.w * s2);
            *cell = o;
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if ((c + 1) % stage_boxes == 0 || c == nch - 1) mbar_arrive(&S->dq_empty[c / stage_boxes]);   // stage drained
            tma_store_3d(&tmap_dq, box, c * 32, 0, (int)p);
            bulk_commit_group();
          }
          __syncwarp();
        }
      }
      if (lane == 0) bulk_wait_group<0>();
    }
  }

#ifdef MMB200_ENABLE_PROF
  if (prof && warp == 1 && lane == 0) {   // per CTA: cycles and nanoseconds from start to the end of the UMMA role
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    prof[304 + 2 * blockIdx.x] = clock64() - t_start;
    prof[304 + 2 * blockIdx.x + 1] = (long long)(ns - t_ns);
  }
  if (prof && blockIdx.x == 0 && lane == 0) {
    const int role = warp == 0 ? 0 : warp == 1 ? 1 : warp == kDdWarp0 ? 2 : warp == kGWarp0 ? 3 : warp == kGWarp0 + 4 ? 4 : warp == kDqWarp ? 5 : warp == 3 ? 6 : -1;
    if (role >= 0) {
      for (int i = 0; i < 6; ++i) prof[role * 7 + i] = pc[i];
      prof[role * 7 + 6] = clock64() - t_start;
    }
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}
#undef KPB_T

template <int KB, bool GATE>
int launch(const KpParams& P, const DeviceInfo& dev, cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& td,
           const CUtensorMap& tdq, const CUtensorMap& tdd) {
  constexpr int KBP = (KB + 3) & ~3;
  const size_t fixed = 1024 + 2 * (size_t)kG2Bytes + (size_t)kDqSlots * kQBoxBytes + sizeof(BwShared<KBP>);
  int stage_boxes = kStageBoxes;
  if (const char* e = getenv("MMB200_KPB_BOXES")) stage_boxes = std::max(1, std::min(kStageBoxes, atoi(e)));
  const size_t stage_bytes = (size_t)stage_boxes * (kBoxBytes + kQBoxBytes);
  int n_stages = std::min<int>(kMaxStages, (int)(((size_t)dev.max_smem_optin - fixed) / stage_bytes));
  if (const char* e = getenv("MMB200_KPB_STAGES")) n_stages = std::max(1, std::min(n_stages, atoi(e)));
  const int mn_sbo = 512;
  if (n_stages < 2) {
    set_error("kernel_pool backward tcgen05: shared-memory plan does not fit");
    return MMB200_ERR_UNSUPPORTED;
  }
  const size_t smem = fixed + (size_t)n_stages * stage_bytes;
  static bool attr_set[64] = {};   // per instantiation and device (the attribute is per device)
  const int di = dev.device & 63;
  if (!attr_set[di]) {
    MMB_CHECK_CUDA(cudaFuncSetAttribute(kernel_pool_bwd_tc_kernel<KB, GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((size_t)dev.max_smem_optin)));
    attr_set[di] = true;
  }
  const int grid = (int)std::min<int64_t>(dev.sm_count, P.B);
  long long* prof = nullptr;
#ifdef MMB200_ENABLE_PROF   // debugging builds only: cudaMalloc + sync in the launch path
  const bool do_prof = getenv("MMB200_KPB_PROF") != nullptr;
  if (do_prof) {
    MMB_CHECK_CUDA(cudaMalloc(&prof, 704 * sizeof(long long)));
    MMB_CHECK_CUDA(cudaMemset(prof, 0, 704 * sizeof(long long)));
  }
#endif
  kernel_pool_bwd_tc_kernel<KB, GATE><<<grid, kThreads, smem, stream>>>(tq, td, tdq, tdd, P, n_stages, stage_boxes, mn_sbo, prof);
  MMB_CHECK_CUDA(cudaGetLastError());
#ifdef MMB200_ENABLE_PROF
  if (do_prof) {
    long long h[704];
    MMB_CHECK_CUDA(cudaStreamSynchronize(stream));
    MMB_CHECK_CUDA(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
    MMB_CHECK_CUDA(cudaFree(prof));
    fprintf(stderr,
            "kpb_prof cycles (CTA 0): tma total %lld wait_q_empty %lld wait_raw_empty %lld | mma total %lld wait_q_full %lld wait_dq_empty %lld "
            "wait_g_full %lld wait_raw_full %lld wait_acc_empty %lld | dd-epi total %lld wait_acc_full %lld bar %lld wait_store_read %lld | "
            "G0 total %lld wait_coef %lld wait_g_empty %lld compute %lld | G1 total %lld wait_coef %lld wait_g_empty %lld compute %lld | "
            "dq total %lld wait_dq_full %lld wait_store_read %lld | store agent total %lld wait_ready %lld wait_read %lld\n",
            h[6], h[0], h[1], h[13], h[7], h[8], h[9], h[10], h[11], h[20], h[14], h[15], h[16], h[27], h[21], h[22], h[23], h[34], h[28],
            h[29], h[30], h[41], h[35], h[36], h[48], h[42], h[43]);
    {
      long long cmax = 0, csum = 0, nmax = 0, nsum = 0;
      for (int b = 0; b < grid; ++b) {
        cmax = std::max(cmax, h[304 + 2 * b]); csum += h[304 + 2 * b];
        nmax = std::max(nmax, h[305 + 2 * b]); nsum += h[305 + 2 * b];
      }
      fprintf(stderr, "kpb_prof per CTA: cycles max %lld mean %lld | ns max %lld mean %lld | clock %.3f GHz\n", cmax, csum / grid, nmax, nsum / grid,
              (double)csum / (double)nsum);
    }
    fprintf(stderr, "kpb_trace tile: MMA arrives | MMA has G | MMA tile issued | G before g_empty | G compute start | G compute end | coef ready (cycles since kernel start of CTA 0's first event)\n");
    {
      long long t0 = h[176 + 6];
      for (int n = 0; n < 16; ++n) {
        const long long* e = h + 176 + 8 * n;
        fprintf(stderr, "  tile %2d: %7lld | %7lld | %7lld | %7lld | %7lld | %7lld | %7lld\n", n, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0,
                e[5] - t0, e[6] ? e[6] - t0 : 0);
      }
    }
    fprintf(stderr, "kpb_trace stage: load_issue | +to raw_full seen by MMA | +MMA issued | +acc_full seen by epilogue | +epilogue done | +agent sees | +store read done (cycles; first column relative to stage 16's issue)\n");
    for (int n = 0; n < 16; ++n) {
      const long long* e = h + 64 + 7 * n;
      fprintf(stderr, "  %2d: %7lld | %6lld | %6lld | %6lld | %6lld | %6lld | %6lld\n", 16 + n, e[0] - h[64], e[1] - e[0], e[2] - e[1], e[3] - e[2],
              e[4] - e[3], e[5] - e[4], e[6] - e[5]);
    }
  }
#endif
  return MMB200_OK;
}

}  // namespace

int kernel_pool_bwd_tc(const KpParams& P, const DeviceInfo& dev, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (P.saved == nullptr || P.Lq > 32 || P.K > 32 || P.D % 4 != 0 || P.D > kMaxD || (P.grad_gate != nullptr && P.gate == nullptr))
    return MMB200_OK;
  if (((reinterpret_cast<uintptr_t>(P.grad_q) | reinterpret_cast<uintptr_t>(P.grad_d) | reinterpret_cast<uintptr_t>(P.saved)) & 15) != 0)
    return MMB200_OK;
  CUtensorMap tq, td, tdq, tdd;
  {
    const uint64_t dims[3] = {(uint64_t)P.D, (uint64_t)P.Lq, (uint64_t)P.B};
    const uint64_t strides[2] = {(uint64_t)P.D * 4, (uint64_t)P.Lq * P.D * 4};
    const uint32_t box[3] = {32, 32, 1};
    if (int rc = encode_tensor_map(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
      return rc;
    if (int rc = encode_tensor_map(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.grad_q, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE))
      return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)P.D, (uint64_t)P.Ld, (uint64_t)P.B};
    const uint64_t strides[2] = {(uint64_t)P.D * 4, (uint64_t)P.Ld * P.D * 4};
    const uint32_t box[3] = {32, 128, 1};
    if (int rc = encode_tensor_map(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.d, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return rc;
    if (int rc = encode_tensor_map(&tdd, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.grad_d, dims, strides, box,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE))
      return rc;
  }
  *handled = true;
  if (P.gate) {   // TK-Sparse: gated activations, d loss / d gate
    if (P.K == 11) return launch<11, true>(P, dev, stream, tq, td, tdq, tdd);
    if (P.K == 21) return launch<21, true>(P, dev, stream, tq, td, tdq, tdd);
    if (P.K <= 12) return launch<12, true>(P, dev, stream, tq, td, tdq, tdd);
    if (P.K <= 24) return launch<24, true>(P, dev, stream, tq, td, tdq, tdd);
    return launch<32, true>(P, dev, stream, tq, td, tdq, tdd);
  }
  if (P.K == 11) return launch<11, false>(P, dev, stream, tq, td, tdq, tdd);
  if (P.K == 21) return launch<21, false>(P, dev, stream, tq, td, tdq, tdd);
  if (P.K <= 12) return launch<12, false>(P, dev, stream, tq, td, tdq, tdd);
  if (P.K <= 24) return launch<24, false>(P, dev, stream, tq, td, tdq, tdd);
  return launch<32, false>(P, dev, stream, tq, td, tdq, tdd);
}

}  // namespace mmb
