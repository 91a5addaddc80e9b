// Fused stem: conv 7x7/2 (+ folded BN) + ReLU + max-pool 3x3/2 on tcgen05 / TMEM fed by TMA  (VQA_OP_STEM_POOL).
// Reference: models/cnn_backbone.py:349-354 (stem Sequential: Conv2d(3,64,7,2,3), BatchNorm2d, ReLU, MaxPool2d(3,2,1)).
//
// Input: the phase-packed image written by ingest_kernel -- one 32-byte row (2x2 pixels x {R,G,B,1|0}) per phase-pixel
// on a padded-flat grid of pitch P = W+2 -- so the 7x7/2 convolution is a 4x4/1 convolution over 16 "channels" and a
// tap (ia, ib) is the row shift (ia-2)*P + (ib-2) of the SAME 2-D tensor (see program.py).
//
// What bounds this layer is the shared-memory port, not the tensor pipe: a 128x64x16 SS-mode MMA reads 6 KB of operands
// for 32 cycles of math, and the first version of the fused stem (gemm_tap_kernel EPI 4) also recomputed every other conv
// row (3 conv rows per pooled row) and pooled a 48 KB tile through shared memory.  This kernel is built around three facts:
//   * ONE M tile is one conv row (128 accumulator rows starting at pixel (2i', 0); P <= 128), and the MMA is N = 128 wide:
//     columns 0..63 are conv row 2i', columns 64..127 conv row 2i'+1.  Both read the same A window -- row 2i'+1 with its
//     taps moved down by one vertical position -- so the weight matrix is [128, 5x4x16] with a zero block at each end
//     (block 0 has no ia' = 4, block 1 no ia' = 0).  20 MMAs of 128x128x16 (8 KB of operand reads per 64 math cycles)
//     replace 2 x 16 MMAs of 128x64x16 (6 KB per 32).
//   * a thread of the epilogue owns one pixel column (TMEM lane) and gets conv rows 2i' and 2i'+1 of that column in its
//     own registers; conv row 2i'-1 is what the same thread held one tile earlier.  A CTA therefore walks RUNS of
//     consecutive pooled rows of one image and carries the previous row in registers: the vertical 3-max costs no memory
//     traffic and no row is computed twice (a run that does not start at the image top begins with one primer tile that
//     only produces the carried row).
//   * only the horizontal 3-max goes through shared memory (16 KB tile, double-buffered: one named barrier per tile).
// The un-pooled 112x112x64 map never exists in HBM.  The folded-BN bias rides in K (ingest writes a 1 column, the centre
// tap's weights there hold the bias as a bf16 hi + lo pair), so the epilogue is ReLU + pack + max.
//
// CTA = 10 warps, one CTA per SM: warp 0 TMA producer (weights once, then one A window per tile), warp 1 MMA issuer
// (one elected thread), warps 2..9 epilogue (lane quadrant x channel half).  Accumulators are double-buffered in TMEM.
#include <cstdio>

#include "common.cuh"

int vqa_encode_2d(CUtensorMap* map, bool tf32, uint64_t base, int rows, int cols, int ld, int box_rows, int row_bytes,
                  const char* what);
int vqa_num_sms(int device);
uint32_t vqa_make_idesc(bool tf32, bool f16, int n, int m);

namespace {

constexpr int kTaps = 20;                    // 5 vertical x 4 horizontal phase-pixel offsets
constexpr int kThreads = 32 * 10;
constexpr int kMaxSlots = 6;
constexpr int kWTile = 128 * 32;             // one tap of the weight matrix: 128 rows x 16 bf16
constexpr int kWBytes = kTaps * kWTile;      // 80 KB, resident for the CTA's lifetime
constexpr int kPoolTile = 128 * 128;         // 128 pixel columns x 64 bf16
constexpr int kAccCols = 128;

struct StemParams {
  int n_runs, run_len, runs_per_img;
  int P, rpi;                // conv grid (= phase-pixel grid): pitch, rows per image
  int Ho, Wo, Po, rpio;      // pooled grid
  int halo_lo;               // window rows before the tile (2P + 2)
  int box_rows, nboxes, a_slot_bytes, a_tx_bytes, a_slots;
  uint32_t idesc, desc_hi;
  uint32_t tap_rel[kTaps];   // row offset of tap t inside the window, in 16-byte units
  uint4* out;
  long long* dbg;            // optional: clock64() timeline of CTA 0 (profiling aid)
};

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                 const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + kWBytes;
  uint8_t* smem_pool = smem_a + p.a_slots * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_pool + 2 * kPoolTile);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxSlots;
  uint64_t* w_full = a_empty + kMaxSlots;
  uint64_t* acc_full = w_full + 1;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool timed = p.dbg != nullptr && blockIdx.x == 0;
  if (timed && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    p.dbg[22] = static_cast<long long>(ns);
    p.dbg[0] = clock64();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
  }
  if (warp == 1) {
    if (lane < kMaxSlots) { mbar_init(&a_full[lane], 1); mbar_init(&a_empty[lane], 1); }
    if (lane == 8) mbar_init(w_full, 1);
    if (lane >= 16 && lane < 18) { mbar_init(&acc_full[lane - 16], 1); mbar_init(&acc_empty[lane - 16], 8); }
    __syncwarp();
    if (lane == 0) mbar_fence_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 2 * kAccCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) pdl_launch_dependents();
  if (timed && threadIdx.x == 0) p.dbg[1] = clock64();

  const int run_len = p.run_len, rpr = p.runs_per_img, n_runs = p.n_runs;
  const int step = static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // weights are constants: fetched before the wait on the preceding kernel
      mbar_expect_tx(w_full, kWBytes);
      for (int t = 0; t < kTaps; ++t) tma_load_2d(smem_w + t * kWTile, &mapW, w_full, t * 16, 0);
    }
    __syncwarp();
    pdl_wait();
    int as = 0;
    uint32_t aph = 0;
    bool first = true;
    for (int run = blockIdx.x; run < n_runs; run += step) {
      const int img = run / rpr, i0 = (run - img * rpr) * run_len;
      for (int k = (i0 == 0 ? 1 : 0); k <= run_len; ++k) {
        const int row0 = img * p.rpi + 2 * (i0 + k - 1) * p.P - p.halo_lo;
        mbar_wait(&a_empty[as], aph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&a_full[as], static_cast<uint32_t>(p.a_tx_bytes));
          uint8_t* dst = smem_a + as * p.a_slot_bytes;
          for (int b = 0; b < p.nboxes; ++b) tma_load_2d(dst + b * p.box_rows * 32, &mapA, &a_full[as], 0, row0 + b * p.box_rows);
        }
        __syncwarp();
        if (first && timed && lane == 0) p.dbg[2] = clock64();
        first = false;
        if (++as == p.a_slots) { as = 0; aph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (elect_one()) {
      const uint32_t a_slot_lo = static_cast<uint32_t>(p.a_slot_bytes) >> 4;
      const uint32_t a_base_lo = smem_u32(smem_a) >> 4;
      const uint32_t w_base_lo = smem_u32(smem_w) >> 4;
      const uint32_t idesc = p.idesc;
      const uint64_t kDescHi = static_cast<uint64_t>(p.desc_hi) << 32;
      const int a_slots = p.a_slots;
      uint32_t rel[kTaps];
#pragma unroll
      for (int t = 0; t < kTaps; ++t) rel[t] = p.tap_rel[t];
      int as = 0, acc = 0;
      uint32_t aph = 0, accph = 0;
      bool first = true;
      long long w_acc = 0, w_a = 0;
      mbar_wait(w_full, 0);
      for (int run = blockIdx.x; run < n_runs; run += step) {
        const int img = run / rpr, i0 = (run - img * rpr) * run_len;
        for (int k = (i0 == 0 ? 1 : 0); k <= run_len; ++k) {
          long long t0 = timed ? clock64() : 0;
          mbar_wait(&acc_empty[acc], accph ^ 1u);
          long long t1 = timed ? clock64() : 0;
          mbar_wait(&a_full[as], aph);
          if (timed) { const long long t2 = clock64(); w_acc += t1 - t0; w_a += t2 - t1; }
          tc_fence_after();
          if (first && timed) p.dbg[3] = clock64();
          const uint32_t d = tmem_base + acc * kAccCols;
          const uint32_t a_lo = a_base_lo + as * a_slot_lo;
#pragma unroll
          for (int t = 0; t < kTaps; ++t)
            umma_f16(d, kDescHi | (a_lo + rel[t]), kDescHi | (w_base_lo + t * (kWTile >> 4)), idesc, t == 0 ? 0u : 1u);
          umma_commit(&a_empty[as]);
          umma_commit(&acc_full[acc]);
          if (first && timed) p.dbg[4] = clock64();
          first = false;
          if (++as == a_slots) { as = 0; aph ^= 1u; }
          if (++acc == 2) { acc = 0; accph ^= 1u; }
        }
      }
      if (timed) { p.dbg[9] = clock64(); p.dbg[18] = w_acc; p.dbg[19] = w_a; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read = 32 pixel columns
    const int half = ew >> 2;                  // which 32 of the 64 channels
    const int tid = threadIdx.x - 64;
    const int px = quad * 32 + lane;           // pixel column of this thread
    const int Po = p.Po, Wo = p.Wo, Ho = p.Ho;
    uint4* const out = p.out;
    uint32_t prev[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) prev[k] = 0u;
    int acc = 0, buf = 0;
    uint32_t accph = 0;
    bool first = true;
    long long w_full_acc = 0;
    pdl_wait();                                // the output grid may still be read by a predecessor
    for (int run = blockIdx.x; run < n_runs; run += step) {
      const int img = run / rpr, i0 = (run - img * rpr) * run_len;
      if (i0 == 0) {                           // conv row -1 is padding: post-ReLU values are >= 0, so 0 is neutral
#pragma unroll
        for (int k = 0; k < 16; ++k) prev[k] = 0u;
      }
      for (int k = (i0 == 0 ? 1 : 0); k <= run_len; ++k) {
        const int ip = i0 + k - 1;             // pooled row of this tile (the primer tile k = 0 produces none)
        const long long t0 = (timed && warp == 2) ? clock64() : 0;
        mbar_wait(&acc_full[acc], accph);
        if (timed && warp == 2) w_full_acc += clock64() - t0;
        tc_fence_after();
        if (first && timed && warp == 2 && lane == 0) p.dbg[5] = clock64();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols + half * 32;
        uint32_t va[32], vb[32];
        __syncwarp();                          // tcgen05.ld is .sync.aligned
        if (k > 0) tmem_ld32(taddr, va);
        tmem_ld32(taddr + 64, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);     // TMEM stage drained: the MMAs of the tile after next may start
        if (++acc == 2) { acc = 0; accph ^= 1u; }
        uint32_t b[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) b[c] = pack_relu_bf16x2(__uint_as_float(vb[2 * c]), __uint_as_float(vb[2 * c + 1]));
        if (k == 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c) prev[c] = b[c];
          continue;
        }
        uint32_t v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint32_t a = pack_relu_bf16x2(__uint_as_float(va[2 * c]), __uint_as_float(va[2 * c + 1]));
          v[c] = max_bf16x2(max_bf16x2(prev[c], a), b[c]);
          prev[c] = b[c];
        }
        uint8_t* const tile = smem_pool + buf * kPoolTile;   // [pixel column][64 bf16], 16-byte chunks XOR-swizzled by column
        uint8_t* const dst = tile + px * 128;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          *reinterpret_cast<uint4*>(dst + (((half * 4 + c4) ^ (px & 7)) << 4)) = make_uint4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
        asm volatile("bar.sync 1, 256;" ::: "memory");       // the vertical maxima of all 128 columns are in the tile
        // ---- horizontal 3-max + store: item = (pooled column j', channel octet cg); the other buffer is written by the
        // next tile only after every thread has passed the next barrier, i.e. finished this pass
        const size_t orow = static_cast<size_t>(img) * p.rpio + static_cast<size_t>(ip) * Po;
        for (int item = tid; item < Po * 8; item += 256) {
          const int jp = item >> 3, cg = item & 7;
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (jp < Wo) {
#pragma unroll
            for (int dc = -1; dc <= 1; ++dc) {
              const int c = 2 * jp + dc;
              if (c < 0) continue;                           // column -1 is padding (c <= W-1 always holds)
              const uint4 q = *reinterpret_cast<const uint4*>(tile + c * 128 + ((cg ^ (c & 7)) << 4));
              o.x = max_bf16x2(o.x, q.x); o.y = max_bf16x2(o.y, q.y); o.z = max_bf16x2(o.z, q.z); o.w = max_bf16x2(o.w, q.w);
            }
          }
          out[(orow + jp) * 8 + cg] = o;
          if (ip == Ho - 1) out[(orow + Po + jp) * 8 + cg] = make_uint4(0u, 0u, 0u, 0u);   // the image's zero pad row
        }
        buf ^= 1;
        if (first && timed && warp == 2 && lane == 0) p.dbg[6] = clock64();
        first = false;
      }
    }
    if (timed && warp == 2 && lane == 0) { p.dbg[7] = clock64(); p.dbg[21] = w_full_acc; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * kAccCols);
  if (timed && threadIdx.x == 0) {
    p.dbg[8] = clock64();
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    p.dbg[23] = static_cast<long long>(ns);
  }
}

struct StemLaunch {
  CUtensorMap mapA, mapW;
  StemParams prm;
  dim3 grid;
  size_t smem;
};

}  // namespace

int stem_launch_bytes() { return static_cast<int>(sizeof(StemLaunch)); }

int stem_prepare(const VqaOp& op, void* storage, int device) {
  StemLaunch* L = new (storage) StemLaunch();
  StemParams& p = L->prm;
  const int32_t* I = op.i;
  const int B = I[STEM_POOL_I_B], H = I[STEM_POOL_I_H], W = I[STEM_POOL_I_W];
  p.P = I[STEM_POOL_I_P];
  p.rpi = I[STEM_POOL_I_RPI];
  p.Ho = I[STEM_POOL_I_Ho]; p.Wo = I[STEM_POOL_I_Wo]; p.Po = I[STEM_POOL_I_Po]; p.rpio = I[STEM_POOL_I_RPIo];
  p.run_len = I[STEM_POOL_I_run_len];
  VQA_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, VQA_E_INVALID, "stem_pool: bad image geometry");
  VQA_REQUIRE(p.P >= W + 2 && p.P <= 128 && p.rpi >= (H + 2) * p.P, VQA_E_INVALID,
              "stem_pool: the conv grid needs two pad columns / rows and a pitch of at most 128");
  VQA_REQUIRE(p.Ho * 2 == H && p.Wo * 2 == W && p.Po >= p.Wo + 1 && p.rpio >= (p.Ho + 1) * p.Po, VQA_E_INVALID,
              "stem_pool: bad pooled grid");
  VQA_REQUIRE(p.run_len >= 1 && p.Ho % p.run_len == 0, VQA_E_INVALID, "stem_pool: run_len must divide the pooled height");
  p.runs_per_img = p.Ho / p.run_len;
  p.n_runs = B * p.runs_per_img;
  p.halo_lo = 2 * p.P + 2;
  const int win = 128 + 4 * p.P + 3;           // taps reach from -(2P+2) to +(2P+1) around the tile
  p.nboxes = (win + 255) / 256;
  p.box_rows = ((win + p.nboxes - 1) / p.nboxes + 7) / 8 * 8;
  p.a_tx_bytes = p.nboxes * p.box_rows * 32;
  p.a_slot_bytes = (p.a_tx_bytes + 1023) / 1024 * 1024;
  for (int ia = 0; ia < 5; ++ia)
    for (int ib = 0; ib < 4; ++ib) p.tap_rel[ia * 4 + ib] = static_cast<uint32_t>(p.halo_lo + (ia - 2) * p.P + (ib - 2)) * 2u;
  p.idesc = vqa_make_idesc(false, false, 128, 128);
  p.desc_hi = static_cast<uint32_t>((8 * 32) >> 4) | (1u << 14) | (6u << 29);   // SBO = 8 rows x 32 B, version 1, SWIZZLE_32B
  const int fixed = 1024 + kWBytes + 2 * kPoolTile + 8 * (2 * kMaxSlots + 1 + 4) + 16;
  int slots = (227 * 1024 - fixed) / p.a_slot_bytes;
  p.a_slots = slots > kMaxSlots ? kMaxSlots : slots;
  VQA_REQUIRE(p.a_slots >= 2, VQA_E_INVALID, "stem_pool: shared memory budget exceeded");
  L->smem = static_cast<size_t>(fixed) + static_cast<size_t>(p.a_slots) * p.a_slot_bytes;
  VQA_REQUIRE(!(op.p[STEM_POOL_P_a] & VQA_EXT_TAG) && !(op.p[STEM_POOL_P_w] & VQA_EXT_TAG) && !(op.p[STEM_POOL_P_out] & VQA_EXT_TAG),
              VQA_E_INVALID, "stem_pool: operands must be arena buffers");
  VQA_REQUIRE(op.p[STEM_POOL_P_out] != 0 && (op.p[STEM_POOL_P_out] & 15) == 0, VQA_E_ALIGN, "stem_pool: output must be 16-byte aligned");
  int rc = vqa_encode_2d(&L->mapA, false, op.p[STEM_POOL_P_a], I[STEM_POOL_I_a_rows], 16, 16, p.box_rows, 32, "stem_pool A");
  if (rc) return rc;
  rc = vqa_encode_2d(&L->mapW, false, op.p[STEM_POOL_P_w], 128, kTaps * 16, kTaps * 16, 128, 32, "stem_pool W");
  if (rc) return rc;
  p.out = reinterpret_cast<uint4*>(op.p[STEM_POOL_P_out]);
  p.dbg = reinterpret_cast<long long*>(op.p[STEM_POOL_P_dbg]);
  int sms = vqa_num_sms(device);
  if (I[STEM_POOL_I_max_ctas] > 0 && I[STEM_POOL_I_max_ctas] < sms) sms = I[STEM_POOL_I_max_ctas];
  int ctas = p.n_runs < sms ? p.n_runs : sms;
  if (p.n_runs > sms) {   // the fewest CTAs that finish in the same number of rounds
    const int rounds = (p.n_runs + sms - 1) / sms;
    ctas = (p.n_runs + rounds - 1) / rounds;
  }
  L->grid = dim3(ctas, 1, 1);
  VQA_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(&stem_pool_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
  return VQA_OK;
}

int stem_run(const void* storage, const uint64_t*, int, cudaStream_t stream) {
  const StemLaunch* L = reinterpret_cast<const StemLaunch*>(storage);
  VQA_CUDA_OK(vqa_launch(stem_pool_kernel, L->grid, dim3(kThreads), L->smem, stream, L->mapA, L->mapW, L->prm));
  VQA_LAUNCH_OK("stem_pool_kernel");
  return VQA_OK;
}
