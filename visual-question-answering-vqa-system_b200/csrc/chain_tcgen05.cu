// Fused post-attention chain of one transformer layer on tcgen05 / TMEM  (VQA_OP_MLP_CHAIN).
//
//   x1   = xres + ctx W_o^T                       attention output projection + residual
//   x2   = x1 + relu(LN(x1) W_1^T + b_1) W_2^T + b_2                    LayerNorm + feed-forward + residual
//   y    = LN'(x2) W_n^T                          optional: the NEXT block's LayerNorm + QKV / query projection
//
// Reference: models/text_encoder.py:373-399 (TransformerEncoderLayer: W_o + residual, norm2, ffn, residual; the next
// layer's norm1 + W_q/W_k/W_v) and models/cross_attention.py:265-299 (CrossAttentionLayer: W_o + residual, norm_ffn, ffn,
// residual; the next layer's norm_query + W_q).  The unfused path runs this as six launches (Linear, LayerNorm, Linear,
// Linear, LayerNorm, Linear) of 5-14 us each for ~2 GFLOP: launch- and latency-bound.  Here a CTA owns 128 token rows for
// the whole chain:
//   * the residual stream lives in TMEM: acc0 (256 fp32 columns) receives ctx W_o^T, the epilogue adds xres IN PLACE
//     (tcgen05.ld / tcgen05.st), and the second FFN GEMM accumulates onto it -- both residual adds cost nothing;
//   * a row of the accumulator is one thread's (TMEM lane = row), so LayerNorm statistics are thread-local sums over
//     the row (two warps share a row's columns and exchange two floats through shared memory);
//   * LayerNorm / ReLU outputs go straight into the K-major SWIZZLE_128B operand tiles of the next GEMM in shared
//     memory (fp16, fence.proxy.async) -- the 1024-wide hidden activation never leaves the SM: it is produced in
//     128-column chunks (acc1, double-buffered) and consumed chunk by chunk as K slices of the second GEMM;
//   * all weights (1.5 MB per layer) stream through one TMA ring in the order the MMA thread consumes them; they
//     are constants, so the ring fills before griddepcontrol.wait.  Every CTA needs the SAME tiles in the same order and
//     an SM pulls only 40-48 B/clk out of L2, so the CTAs of a cluster (CS = 2 or 4 row tiles) share the stream: each
//     loads 1/CS of every tile and multicasts it into all CS rings (cp.async.bulk.tensor ... .multicast::cluster); a ring
//     slot is free when all CS MMA threads have consumed it (tcgen05.commit ... .multicast::cluster onto every CTA's
//     empty barrier).
// Operands are fp16 with fp32 accumulation, like the unfused Linears of the throughput mode (program.py::linear).
//
// CTA = 10 warps: warp 0 weight producer, warp 1 MMA issuer (one elected thread; also loads the ctx tile),
// warps 2..9 epilogue (TMEM lane quadrant x column half).  TMEM: acc0 = columns 0..255, acc1 = 2 x 128 columns.
#include <cstdio>

#include "common.cuh"

int vqa_encode_2d(CUtensorMap* map, bool tf32, uint64_t base, int rows, int cols, int ld, int box_rows, int row_bytes,
                  const char* what);
int vqa_encode_box32f(CUtensorMap* map, uint64_t base, int rows, int cols, int ld, const char* what);
uint32_t vqa_make_idesc(bool tf32, bool f16, int n, int m);

namespace {

constexpr int kD = 256;
constexpr int kThreads = 32 * 10;
constexpr int kRing = 5;               // weight tiles in flight
constexpr int kTile = 128 * 128;       // one operand tile: 128 rows x 64 fp16 (128-byte rows, SWIZZLE_128B) = 16 KB
constexpr int kTileLo = kTile >> 4;    // in 16-byte descriptor units
constexpr int kABytes = 4 * kTile;     // A operand of the K = 256 GEMMs
constexpr int kHBytes = 2 * kTile;     // one 128-wide chunk of the hidden activation (K = 128)
constexpr int kMaxF = 1024;
constexpr int kSlot = 32 * 128;        // epilogue box: 32 rows x 32 fp32, SWIZZLE_128B

// shared-memory map (offsets from the 1024-byte aligned base)
constexpr int kOffA = 0;
constexpr int kOffH = kOffA + kABytes;                 // H[2]; doubles as the epilogue's residual / staging slots
constexpr int kOffRing = kOffH + 2 * kHBytes;
constexpr int kOffTab = kOffRing + kRing * kTile;      // b1[kMaxF], b2, ln_g, ln_b, n_g, n_b [256 each]
constexpr int kOffStat = kOffTab + 4 * (kMaxF + 5 * kD);   // [2 halves][128 rows] x {sum, sq}
constexpr int kOffBar = kOffStat + 4 * 4 * 128;
constexpr int kNumBars = 2 * kRing + 16 + 16;
constexpr int kSmemBytes = 1024 + kOffBar + 8 * kNumBars + 16;
static_assert(kSmemBytes <= 227 * 1024, "chain kernel: shared memory budget");
static_assert(8 * 2 * kSlot <= 2 * kHBytes, "epilogue slots must fit the hidden-activation buffers");

struct ChainParams {
  int T, m_tiles, F, Nn;               // rows, 128-row tiles (padded to the cluster size), hidden width, follow-up projection width (0 = none)
  int CS;                              // cluster size: CTAs that share the weight stream
  const float *b1, *b2, *ln_g, *ln_b, *n_g, *n_b;
  float eps, eps_n;
  uint32_t idesc;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one part of a weight tile into the same ring slot of every CTA in `mask`; each CTA's barrier at this offset gets the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t x, int32_t y, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(mask)
      : "memory");
}
// the barrier at this offset in every CTA of `mask` receives one arrival when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
mlp_chain_kernel(const __grid_constant__ CUtensorMap mapCtx, const __grid_constant__ CUtensorMap mapWo,
                 const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                 const __grid_constant__ CUtensorMap mapWn, const __grid_constant__ CUtensorMap mapXres,
                 const __grid_constant__ CUtensorMap mapXout, const __grid_constant__ CUtensorMap mapY,
                 const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* const s_a = smem + kOffA;
  uint8_t* const s_h = smem + kOffH;
  uint8_t* const s_ring = smem + kOffRing;
  float* const s_b1 = reinterpret_cast<float*>(smem + kOffTab);
  float* const s_b2 = s_b1 + kMaxF;
  float* const s_g = s_b2 + kD;
  float* const s_b = s_g + kD;
  float* const s_ng = s_b + kD;
  float* const s_nb = s_ng + kD;
  float* const s_sum = reinterpret_cast<float*>(smem + kOffStat);   // [2][128]
  float* const s_sq = s_sum + 256;                                  // [2][128]
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* const b_full = bars;
  uint64_t* const b_empty = b_full + kRing;
  uint64_t* const a_full = b_empty + kRing;       // ctx tile landed
  uint64_t* const a_empty = a_full + 1;           // every MMA that reads the A buffer has retired
  uint64_t* const acc0_ready = a_empty + 1;       // ctx W_o^T complete
  uint64_t* const xn_ready = acc0_ready + 1;      // LN(x1) is in the A buffer, x1 is in acc0
  uint64_t* const acc1_full = xn_ready + 1;       // [2]
  uint64_t* const h_ready = acc1_full + 2;        // [2] hidden chunk written (and acc1 stage drained)
  uint64_t* const h_empty = h_ready + 2;          // [2] the MMAs that read the hidden chunk have retired
  uint64_t* const acc0_final = h_empty + 2;       // x2 - b_2 complete
  uint64_t* const xn2_ready = acc0_final + 1;     // LN'(x2) is in the A buffer
  uint64_t* const acc1_free = xn2_ready + 1;      // [2] follow-up projection chunk drained
  uint64_t* const acc0_free = acc1_free + 2;      // the epilogue has read x2 out of acc0
  uint64_t* const res_bar = acc0_free + 1;        // [8 warps][2 slots]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nF = p.F / 128, nN = p.Nn / 128;
  const int CS = p.CS;
  const uint32_t rank = CS > 1 ? cluster_rank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << CS) - 1u);
  const int part_rows = 128 / CS;      // weight-tile rows this CTA fetches for the whole cluster

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapCtx); tma_prefetch_desc(&mapWo); tma_prefetch_desc(&mapW1); tma_prefetch_desc(&mapW2);
    if (nN) { tma_prefetch_desc(&mapWn); tma_prefetch_desc(&mapY); }
    tma_prefetch_desc(&mapXres); tma_prefetch_desc(&mapXout);
  }
  if (warp == 1) {
    if (lane < kRing) { mbar_init(&b_full[lane], 1); mbar_init(&b_empty[lane], static_cast<uint32_t>(CS)); }
    if (lane == 8) {
      mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(acc0_ready, 1); mbar_init(xn_ready, 8);
      mbar_init(acc0_final, 1); mbar_init(xn2_ready, 8); mbar_init(acc0_free, 8);
    }
    if (lane >= 10 && lane < 12) {
      mbar_init(&acc1_full[lane - 10], 1); mbar_init(&h_ready[lane - 10], 8); mbar_init(&h_empty[lane - 10], 1);
      mbar_init(&acc1_free[lane - 10], 8);
    }
    if (lane >= 16) mbar_init(&res_bar[lane - 16], 1);
    __syncwarp();
    if (lane == 0) mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // bias / LayerNorm tables (constants)
  for (int i = threadIdx.x; i < p.F; i += kThreads) s_b1[i] = __ldg(p.b1 + i);
  for (int i = threadIdx.x; i < kD; i += kThreads) {
    s_b2[i] = __ldg(p.b2 + i); s_g[i] = __ldg(p.ln_g + i); s_b[i] = __ldg(p.ln_b + i);
    s_ng[i] = nN ? __ldg(p.n_g + i) : 0.f; s_nb[i] = nN ? __ldg(p.n_b + i) : 0.f;
  }
  tc_fence_before();
  if (CS > 1) cluster_sync(); else __syncthreads();   // peers multicast into this CTA's ring and signal its barriers
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) pdl_launch_dependents();

  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp == 0) {
    // ===================== weight producer: every tile of every GEMM, in the order the MMA thread consumes them =========
    int ring = 0;
    uint32_t rph = 0;
    auto put = [&](const CUtensorMap* m, int n0, int k0) {
      mbar_wait(&b_empty[ring], rph ^ 1u);               // every CTA of the cluster has consumed this slot
      if (elect_one()) {
        mbar_expect_tx(&b_full[ring], kTile);             // the whole tile: this CTA's part plus the peers' multicasts
        if (CS > 1) tma_load_2d_mc(s_ring + ring * kTile + rank * part_rows * 128, m, &b_full[ring], k0, n0 + rank * part_rows, cmask);
        else tma_load_2d(s_ring + ring * kTile, m, &b_full[ring], k0, n0);
      }
      __syncwarp();
      if (++ring == kRing) { ring = 0; rph ^= 1u; }
    };
    auto g2 = [&](int j) { for (int kc = 0; kc < 4; ++kc) put(&mapW1, j * 128, kc * 64); };
    auto g3 = [&](int j) { for (int kc = 0; kc < 2; ++kc) for (int nh = 0; nh < 2; ++nh) put(&mapW2, nh * 128, j * 128 + kc * 64); };
    for (int tile = first_tile; tile < p.m_tiles; tile += tile_step) {
      for (int kc = 0; kc < 4; ++kc) for (int nh = 0; nh < 2; ++nh) put(&mapWo, nh * 128, kc * 64);
      g2(0);
      if (nF > 1) g2(1);
      for (int j = 0; j < nF; ++j) { g3(j); if (j + 2 < nF) g2(j + 2); }
      for (int n = 0; n < nN; ++n) for (int kc = 0; kc < 4; ++kc) put(&mapWn, n * 128, kc * 64);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint64_t kDescHi = umma_desc_sw128_hi() << 32;
      const uint32_t idesc = p.idesc;
      const uint32_t a_lo = smem_u32(s_a) >> 4, h_lo = smem_u32(s_h) >> 4, ring_lo = smem_u32(s_ring) >> 4;
      const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256;
      int ring = 0;
      uint32_t rph = 0, ph_a = 0, ph_ae = 0, ph_xn = 0, ph_xn2 = 0, ph_a0f = 0;
      uint32_t ph_h[2] = {0, 0}, ph_free[2] = {0, 0};
      // 4 K slices of one 64-wide K chunk: D[128 x 128] (+)= A[128 x 64] B[128 x 64]^T
      auto mma_tile = [&](uint32_t d, uint32_t a, bool fresh) {
        mbar_wait(&b_full[ring], rph);
        tc_fence_after();
        const uint32_t b = ring_lo + ring * kTileLo;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d, kDescHi | (a + 2 * k), kDescHi | (b + 2 * k), idesc, (fresh && k == 0) ? 0u : 1u);
        if (CS > 1) umma_commit_mc(&b_empty[ring], cmask); else umma_commit(&b_empty[ring]);
        if (++ring == kRing) { ring = 0; rph ^= 1u; }
      };
      auto g2 = [&](int j) {     // acc1[j & 1] = LN(x1) W_1[j]^T
        for (int kc = 0; kc < 4; ++kc) mma_tile(acc1 + (j & 1) * 128, a_lo + kc * kTileLo, kc == 0);
        umma_commit(&acc1_full[j & 1]);
      };
      auto g3 = [&](int j) {     // acc0 += hidden[j] W_2[:, j]^T
        mbar_wait(&h_ready[j & 1], ph_h[j & 1]);
        ph_h[j & 1] ^= 1u;
        tc_fence_after();
        for (int kc = 0; kc < 2; ++kc)
          for (int nh = 0; nh < 2; ++nh) mma_tile(acc0 + nh * 128, h_lo + (j & 1) * (kHBytes >> 4) + kc * kTileLo, false);
        umma_commit(&h_empty[j & 1]);
      };
      bool first = true;
      for (int tile = first_tile; tile < p.m_tiles; tile += tile_step) {
        const int m0 = tile * 128;
        if (!first) {            // the previous tile's MMAs have read the A buffer, its epilogue has read acc0
          mbar_wait(a_empty, ph_ae); ph_ae ^= 1u;
          mbar_wait(acc0_free, ph_a0f); ph_a0f ^= 1u;
          tc_fence_after();
        } else {
          pdl_wait();            // ctx is written by the preceding attention kernel
        }
        first = false;
        mbar_expect_tx(a_full, kABytes);
        for (int kc = 0; kc < 4; ++kc) tma_load_2d(s_a + kc * kTile, &mapCtx, a_full, kc * 64, m0);
        mbar_wait(a_full, ph_a); ph_a ^= 1u;
        tc_fence_after();
        for (int kc = 0; kc < 4; ++kc)
          for (int nh = 0; nh < 2; ++nh) mma_tile(acc0 + nh * 128, a_lo + kc * kTileLo, kc == 0);
        umma_commit(acc0_ready);
        mbar_wait(xn_ready, ph_xn); ph_xn ^= 1u;
        tc_fence_after();
        g2(0);
        if (nF > 1) g2(1);
        for (int j = 0; j < nF; ++j) { g3(j); if (j + 2 < nF) g2(j + 2); }
        umma_commit(acc0_final);
        if (nN) {
          mbar_wait(xn2_ready, ph_xn2); ph_xn2 ^= 1u;
          tc_fence_after();
          for (int n = 0; n < nN; ++n) {
            if (n >= 2) { mbar_wait(&acc1_free[n & 1], ph_free[n & 1]); ph_free[n & 1] ^= 1u; tc_fence_after(); }
            for (int kc = 0; kc < 4; ++kc) mma_tile(acc1 + (n & 1) * 128, a_lo + kc * kTileLo, kc == 0);
            umma_commit(&acc1_full[n & 1]);
          }
          // the drains of the last two chunks are not waited for by a later chunk: consume their arrivals here so the
          // barrier phases stay aligned for the next tile
          for (int n = (nN >= 2 ? nN - 2 : 0); n < nN; ++n) { mbar_wait(&acc1_free[n & 1], ph_free[n & 1]); ph_free[n & 1] ^= 1u; }
        }
        umma_commit(a_empty);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                   // which half of a GEMM's output columns
    const int r = quad * 32 + lane;             // row of the tile this thread owns
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t acc0 = lane_addr, acc1 = lane_addr + 256;
    uint8_t* const slots = s_h + ew * (2 * kSlot);       // two 4 KB boxes per warp (H is idle whenever they are used)
    uint64_t* const my_res = res_bar + ew * 2;
    const uint32_t swz = lane & 7;                       // SWIZZLE_128B: 16-byte unit u of row `lane` sits at (u ^ swz) << 4
    uint32_t ph_res[2] = {0, 0}, ph_acc1[2] = {0, 0}, ph_he[2] = {0, 0};
    uint32_t ph_a0 = 0, ph_fin = 0;
    const float eps = p.eps, eps_n = p.eps_n;
    pdl_wait();                                          // xres / xout / y belong to the preceding kernels until now

    // LayerNorm of the row held in acc0 (+ optional per-column constant): statistics over this thread's 128 columns, the
    // other half's through shared memory; writes the normalised row as fp16 into the A buffer (K-major, SWIZZLE_128B)
    auto layer_norm_to_a = [&](const float* add, const float* gam, const float* bet, float e, float sum) {
      s_sum[half * 128 + r] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mean = (s_sum[r] + s_sum[128 + r]) * (1.f / 256.f);
      float q = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(acc0 + half * 128 + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float d = __uint_as_float(v[k]) + (add ? add[half * 128 + 32 * c + k] : 0.f) - mean;
          q += d * d;
        }
      }
      s_sq[half * 128 + r] = q;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float rstd = rsqrtf((s_sq[r] + s_sq[128 + r]) * (1.f / 256.f) + e);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(acc0 + half * 128 + 32 * c, v);
        tmem_ld_wait();
        const int col = half * 128 + 32 * c;
        uint8_t* const dst = s_a + (col >> 6) * kTile + r * 128;
        const int u0 = (c & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i0 = 8 * u + 2 * k;
            const float y0 = (__uint_as_float(v[i0]) + (add ? add[col + i0] : 0.f) - mean) * rstd * gam[col + i0] + bet[col + i0];
            const float y1 = (__uint_as_float(v[i0 + 1]) + (add ? add[col + i0 + 1] : 0.f) - mean) * rstd * gam[col + i0 + 1] + bet[col + i0 + 1];
            w[k] = pack_f16x2(y0, y1);
          }
          *reinterpret_cast<uint4*>(dst + (((u0 + u) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    };

    for (int tile = first_tile; tile < p.m_tiles; tile += tile_step) {
      const int m0 = tile * 128, row0 = m0 + quad * 32;
      // ---- E1: x1 = acc0 + xres (written back to acc0), LN(x1) -> A buffer
      auto issue_res = [&](int c) {                      // lane 0: 32 rows x 32 fp32 of the residual into slot c & 1
        mbar_expect_tx(&my_res[c & 1], kSlot);
        tma_load_2d(slots + (c & 1) * kSlot, &mapXres, &my_res[c & 1], half * 128 + 32 * c, row0);
      };
      if (lane == 0) { issue_res(0); issue_res(1); }
      mbar_wait(acc0_ready, ph_a0); ph_a0 ^= 1u;
      tc_fence_after();
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(acc0 + half * 128 + 32 * c, v);
        mbar_wait(&my_res[c & 1], ph_res[c & 1]); ph_res[c & 1] ^= 1u;
        tmem_ld_wait();
        const uint8_t* const rrow = slots + (c & 1) * kSlot + lane * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 q4 = *reinterpret_cast<const float4*>(rrow + ((u ^ swz) << 4));
          const float x0 = __uint_as_float(v[4 * u]) + q4.x, x1 = __uint_as_float(v[4 * u + 1]) + q4.y;
          const float x2 = __uint_as_float(v[4 * u + 2]) + q4.z, x3 = __uint_as_float(v[4 * u + 3]) + q4.w;
          sum += (x0 + x1) + (x2 + x3);
          v[4 * u] = __float_as_uint(x0); v[4 * u + 1] = __float_as_uint(x1);
          v[4 * u + 2] = __float_as_uint(x2); v[4 * u + 3] = __float_as_uint(x3);
        }
        tmem_st32(acc0 + half * 128 + 32 * c, v);
        __syncwarp();                                    // every lane has consumed the slot
        if (lane == 0 && c + 2 < 4) issue_res(c + 2);
      }
      tmem_st_wait();
      layer_norm_to_a(nullptr, s_g, s_b, eps, sum);
      fence_proxy_async();                               // A-buffer writes -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(xn_ready);

      // ---- E2: hidden chunk j = relu(acc1[j & 1] + b_1) -> H[j & 1] (fp16, K-major, SWIZZLE_128B)
#pragma unroll 1
      for (int j = 0; j < nF; ++j) {
        const int b = j & 1;
        mbar_wait(&acc1_full[b], ph_acc1[b]); ph_acc1[b] ^= 1u;
        tc_fence_after();
        if (j >= 2) { mbar_wait(&h_empty[b], ph_he[b]); ph_he[b] ^= 1u; }   // chunk j-2's MMAs have retired
        uint8_t* const dst = s_h + b * kHBytes + half * kTile + r * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(acc1 + b * 128 + half * 64 + 32 * c, v);
          tmem_ld_wait();
          const float* const bias = s_b1 + j * 128 + half * 64 + 32 * c;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i0 = 8 * u + 2 * k;
              w[k] = pack_relu_f16x2(__uint_as_float(v[i0]) + bias[i0], __uint_as_float(v[i0 + 1]) + bias[i0 + 1]);
            }
            *reinterpret_cast<uint4*>(dst + (((4 * c + u) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[b]);
      }
      // the last two h_empty arrivals are not waited for above: consume them so the phases stay aligned
      for (int j = (nF >= 2 ? nF - 2 : 0); j < nF; ++j) { mbar_wait(&h_empty[j & 1], ph_he[j & 1]); ph_he[j & 1] ^= 1u; }

      // ---- E3: x2 = acc0 + b_2 -> xout (fp32, TMA store); LN'(x2) -> A buffer
      mbar_wait(acc0_final, ph_fin); ph_fin ^= 1u;
      tc_fence_after();
      sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(acc0 + half * 128 + 32 * c, v);
        tmem_ld_wait();
        const int col = half * 128 + 32 * c;
        uint8_t* const srow = slots + (c & 1) * kSlot + lane * 128;
        if (lane == 0 && c >= 2) bulk_wait_read0();      // the store issued two chunks ago has drained this slot
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float x0 = __uint_as_float(v[4 * u]) + s_b2[col + 4 * u], x1 = __uint_as_float(v[4 * u + 1]) + s_b2[col + 4 * u + 1];
          const float x2 = __uint_as_float(v[4 * u + 2]) + s_b2[col + 4 * u + 2], x3 = __uint_as_float(v[4 * u + 3]) + s_b2[col + 4 * u + 3];
          sum += (x0 + x1) + (x2 + x3);
          *reinterpret_cast<float4*>(srow + ((u ^ swz) << 4)) = make_float4(x0, x1, x2, x3);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(&mapXout, slots + (c & 1) * kSlot, col, row0); bulk_commit(); }
      }
      if (nN) {
        layer_norm_to_a(s_b2, s_ng, s_nb, eps_n, sum);
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(xn2_ready); mbar_arrive(acc0_free); }
        // ---- E4: follow-up projection chunk n -> y (fp32, TMA store)
#pragma unroll 1
        for (int n = 0; n < nN; ++n) {
          const int b = n & 1;
          mbar_wait(&acc1_full[b], ph_acc1[b]); ph_acc1[b] ^= 1u;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(acc1 + b * 128 + half * 64 + 32 * c, v);
            tmem_ld_wait();
            uint8_t* const srow = slots + c * kSlot + lane * 128;
            if (lane == 0) bulk_wait_read0();            // at most one store is left outstanding per slot pair
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<float4*>(srow + ((u ^ swz) << 4)) =
                  make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&mapY, slots + c * kSlot, n * 128 + half * 64 + 32 * c, row0); bulk_commit(); }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc1_free[b]);
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc0_free);
      }
      if (lane == 0) bulk_wait_read0();                  // the slots are reused for the next tile's residual boxes
      __syncwarp();
    }
    if (lane == 0) bulk_wait_all();                      // outstanding TMA stores complete before the CTA exits
  }

  tc_fence_before();
  if (CS > 1) cluster_sync(); else __syncthreads();     // peers' last multicasts / commits land in this CTA
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

struct ChainLaunch {
  CUtensorMap mapCtx, mapWo, mapW1, mapW2, mapWn, mapXres, mapXout, mapY;
  ChainParams prm;
  dim3 grid;
};

}  // namespace

int chain_launch_bytes() { return static_cast<int>(sizeof(ChainLaunch)); }

int chain_prepare(const VqaOp& op, void* storage, int device) {
  ChainLaunch* L = new (storage) ChainLaunch();
  ChainParams& p = L->prm;
  const int32_t* I = op.i;
  p.T = I[MLP_CHAIN_I_T];
  p.F = I[MLP_CHAIN_I_F];
  p.Nn = I[MLP_CHAIN_I_Nn];
  VQA_REQUIRE(p.T > 0 && I[MLP_CHAIN_I_D] == kD, VQA_E_INVALID, "mlp_chain: D must be 256");
  VQA_REQUIRE(p.F >= 128 && p.F % 128 == 0 && p.F <= kMaxF, VQA_E_INVALID, "mlp_chain: hidden width must be a multiple of 128, <= 1024");
  VQA_REQUIRE(p.Nn >= 0 && p.Nn % 128 == 0, VQA_E_INVALID, "mlp_chain: follow-up width must be a multiple of 128");
  p.m_tiles = (p.T + 127) / 128;
  // measured on B200 (256 pairs, 40 row tiles): clusters of 1 / 2 / 4 give the same step time within noise -- the chain is
  // bound by the shared-memory port (N = 128 MMAs read 8 KB per 64 cycles while the ring is being filled), not by L2 --
  // so the default is no cluster; VQA_CHAIN_CS / the op's CS field select the multicast form
  p.CS = I[MLP_CHAIN_I_CS] > 0 ? I[MLP_CHAIN_I_CS] : 1;
  VQA_REQUIRE(p.CS == 1 || p.CS == 2 || p.CS == 4, VQA_E_INVALID, "mlp_chain: cluster size must be 1, 2 or 4");
  p.m_tiles = (p.m_tiles + p.CS - 1) / p.CS * p.CS;   // phantom tiles keep the cluster's weight stream in lockstep (loads read
                                                       // zeros past T, stores are clipped)
  for (int k = 0; k < MLP_CHAIN_NP; ++k)
    VQA_REQUIRE(!(op.p[k] & VQA_EXT_TAG) && (op.p[k] & 15) == 0, VQA_E_INVALID, "mlp_chain: operands must be 16-byte aligned arena buffers");
  auto P = [&](int k) { return op.p[k]; };
  VQA_REQUIRE(P(MLP_CHAIN_P_ctx) && P(MLP_CHAIN_P_xres) && P(MLP_CHAIN_P_xout) && P(MLP_CHAIN_P_wo) && P(MLP_CHAIN_P_w1) &&
                  P(MLP_CHAIN_P_w2) && P(MLP_CHAIN_P_b1) && P(MLP_CHAIN_P_b2) && P(MLP_CHAIN_P_ln_g) && P(MLP_CHAIN_P_ln_b),
              VQA_E_INVALID, "mlp_chain: null operand");
  VQA_REQUIRE(p.Nn == 0 || (P(MLP_CHAIN_P_wn) && P(MLP_CHAIN_P_n_g) && P(MLP_CHAIN_P_n_b) && P(MLP_CHAIN_P_y)), VQA_E_INVALID,
              "mlp_chain: the follow-up projection needs its LayerNorm, weights and output");
  p.b1 = reinterpret_cast<const float*>(P(MLP_CHAIN_P_b1));
  p.b2 = reinterpret_cast<const float*>(P(MLP_CHAIN_P_b2));
  p.ln_g = reinterpret_cast<const float*>(P(MLP_CHAIN_P_ln_g));
  p.ln_b = reinterpret_cast<const float*>(P(MLP_CHAIN_P_ln_b));
  p.n_g = reinterpret_cast<const float*>(P(MLP_CHAIN_P_n_g));
  p.n_b = reinterpret_cast<const float*>(P(MLP_CHAIN_P_n_b));
  p.eps = op.f[MLP_CHAIN_F_eps];
  p.eps_n = op.f[MLP_CHAIN_F_eps_n];
  p.idesc = vqa_make_idesc(false, true, 128, 128);
  int rc = vqa_encode_2d(&L->mapCtx, false, P(MLP_CHAIN_P_ctx), p.T, kD, kD, 128, 128, "mlp_chain ctx");
  if (rc) return rc;
  const int wbox = 128 / p.CS;     // each CTA of a cluster fetches 1/CS of every weight tile
  rc = vqa_encode_2d(&L->mapWo, false, P(MLP_CHAIN_P_wo), kD, kD, kD, wbox, 128, "mlp_chain W_o");
  if (rc) return rc;
  rc = vqa_encode_2d(&L->mapW1, false, P(MLP_CHAIN_P_w1), p.F, kD, kD, wbox, 128, "mlp_chain W_1");
  if (rc) return rc;
  rc = vqa_encode_2d(&L->mapW2, false, P(MLP_CHAIN_P_w2), kD, p.F, p.F, wbox, 128, "mlp_chain W_2");
  if (rc) return rc;
  L->mapWn = L->mapWo;
  L->mapY = L->mapCtx;
  if (p.Nn) {
    rc = vqa_encode_2d(&L->mapWn, false, P(MLP_CHAIN_P_wn), I[MLP_CHAIN_I_Nn_pad], kD, kD, wbox, 128, "mlp_chain W_n");
    if (rc) return rc;
    rc = vqa_encode_box32f(&L->mapY, P(MLP_CHAIN_P_y), p.T, p.Nn, p.Nn, "mlp_chain y");
    if (rc) return rc;
  }
  rc = vqa_encode_box32f(&L->mapXres, P(MLP_CHAIN_P_xres), p.T, kD, kD, "mlp_chain xres");
  if (rc) return rc;
  rc = vqa_encode_box32f(&L->mapXout, P(MLP_CHAIN_P_xout), p.T, kD, kD, "mlp_chain xout");
  if (rc) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (I[MLP_CHAIN_I_max_ctas] > 0 && I[MLP_CHAIN_I_max_ctas] < sms) sms = I[MLP_CHAIN_I_max_ctas];
  sms = sms / p.CS * p.CS;
  L->grid = dim3(p.m_tiles < sms ? p.m_tiles : (sms > 0 ? sms : p.CS), 1, 1);
  VQA_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(&mlp_chain_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSmemBytes));
  return VQA_OK;
}

int chain_run(const void* storage, const uint64_t*, int, cudaStream_t stream) {
  const ChainLaunch* L = reinterpret_cast<const ChainLaunch*>(storage);
  VQA_CUDA_OK(vqa_launch_cluster(mlp_chain_kernel, L->grid, dim3(kThreads), kSmemBytes, stream, L->prm.CS, L->mapCtx, L->mapWo,
                                 L->mapW1, L->mapW2, L->mapWn, L->mapXres, L->mapXout, L->mapY, L->prm));
  VQA_LAUNCH_OK("mlp_chain_kernel");
  return VQA_OK;
}
