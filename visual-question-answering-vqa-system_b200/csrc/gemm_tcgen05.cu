// Tap-shifted GEMM on tcgen05 / TMEM fed by TMA  (VQA_OP_GEMM), persistent and warp-specialised.
//
// One kernel covers every Conv2d(+BN folded)(+ReLU)(+residual) of the backbone
// (reference models/cnn_backbone.py:164-197, 349-352) and every nn.Linear of the text encoder,
// fusion module and answer head: out[M,N] = sum_g sum_t A_g[m + delta_g - halo + rel_t, :] * W_t^T.
//
// Data layout (see program.py): A is a 2-D K-major tensor (rows = padded-flat pixels or tokens,
// 128-byte K chunks), W is [Npad, Ktot] K-major.  Per K chunk the producer warp TMA-loads ONE
// window of A rows (tile rows + halo on both sides, SWIZZLE_128B) and the MMA thread issues one
// UMMA set per tap whose A descriptor starts `rel` rows into that window -- a 3x3 convolution
// re-uses each input row 9 times out of shared memory instead of re-fetching it from L2.
// (Measured on B200: the UMMA SWIZZLE_128B pattern is a function of the absolute shared-memory
// address, so a start address offset by whole 128-byte rows needs no base_offset correction.)
//
// CTA = 10 warps, one CTA per SM, looping over output tiles (static round-robin):
//   warp 0      TMA producer: A-window ring + weight ring (or all weights resident when they fit)
//   warp 1      MMA issuer: one thread issues tcgen05.mma; tcgen05.commit frees ring slots
//   warps 2..9  epilogue: TMEM -> registers -> bias / residual / ReLU / pad-mask -> global; two
//               warps per TMEM lane quadrant, each taking half of the tile's columns
// Accumulators are double-buffered in TMEM (when 2*MT*BN <= 512 columns) so the epilogue of tile
// i overlaps the MMAs of tile i+1.
#include <cstdio>

#include "common.cuh"

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kChunkBytes = 128;  // one SWIZZLE_128B row: 64 bf16 or 32 fp32 (tf32)
constexpr int kMaxASlots = 8;
constexpr int kMaxBSlots = 12;

struct GemmParams {
  int M, N;
  int MT;            // 128-row sub-tiles per CTA tile (1 or 2); mirrors the kernel's template MT
  int halo;          // window rows before/after the tile
  int box_rows;      // TMA box rows for A
  int nboxes;        // boxes per window (1 or 2)
  int a_slots, b_slots;
  int a_slot_bytes, b_slot_bytes;
  int b_resident;    // all weight chunks of the N tile stay in shared memory for the CTA's lifetime
  int k_chunks;      // Ktot / chunk_elems
  int acc_stages;    // TMEM accumulator double buffering (1 or 2)
  int m_tiles, n_tiles;
  int ngroups;
  int chunk_elems;   // 64 (bf16) / 32 (tf32)
  int is_tf32;
  uint32_t idesc;
  int g_map[VQA_MAX_GROUPS], g_delta[VQA_MAX_GROUPS], g_acol[VQA_MAX_GROUPS], g_chunks[VQA_MAX_GROUPS];
  int g_ntaps[VQA_MAX_GROUPS], g_kbase[VQA_MAX_GROUPS], g_tap0[VQA_MAX_GROUPS];
  int tap_rel[VQA_MAX_TAPS];
  // epilogue
  void* out;
  const float* bias;
  const void* res;
  int ldo, ldr, out_dtype, res_dtype;
  int relu, round_tf32, mask_en, mP, mRPI, mH, mW;
};

// ---- epilogue helpers -----------------------------------------------------------------------
// Residual prefetch for 32 columns of one row: bf16 -> 4 x uint4, fp32 -> 8 x uint4.
struct ResRegs { uint4 q[8]; };

__device__ __forceinline__ void res_prefetch(ResRegs& r, const GemmParams& p, int row, int col, bool row_ok) {
  if (!p.res || !row_ok) return;
  if (p.res_dtype == 0) {
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.res) + static_cast<size_t>(row) * p.ldr + col;
    if (col + 32 <= p.N && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r.q[j] = reinterpret_cast<const uint4*>(src)[j];
    }
  } else {
    const float* src = reinterpret_cast<const float*>(p.res) + static_cast<size_t>(row) * p.ldr + col;
    if (col + 32 <= p.N && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r.q[j] = reinterpret_cast<const uint4*>(src)[j];
    }
  }
}

__device__ __forceinline__ void res_apply(float (&x)[32], const ResRegs& r, const GemmParams& p, int row, int col) {
  if (!p.res) return;
  if (p.res_dtype == 0) {
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.res) + static_cast<size_t>(row) * p.ldr + col;
    if (col + 32 <= p.N && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {r.q[j].x, r.q[j].y, r.q[j].z, r.q[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          x[8 * j + 2 * k] += __uint_as_float(w[k] << 16);
          x[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
        }
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (col + j < p.N) x[j] += __bfloat162float(src[j]);
    }
  } else {
    const float* src = reinterpret_cast<const float*>(p.res) + static_cast<size_t>(row) * p.ldr + col;
    if (col + 32 <= p.N && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[4 * j] += __uint_as_float(r.q[j].x);
        x[4 * j + 1] += __uint_as_float(r.q[j].y);
        x[4 * j + 2] += __uint_as_float(r.q[j].z);
        x[4 * j + 3] += __uint_as_float(r.q[j].w);
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (col + j < p.N) x[j] += src[j];
    }
  }
}

__device__ __forceinline__ void store_row32(const float (&x)[32], const GemmParams& p, int row, int col) {
  if (p.out_dtype == 0) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
    if (col + 32 <= p.N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q;
        q.x = pack_bf16x2(x[8 * j], x[8 * j + 1]);
        q.y = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
        q.z = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]);
        q.w = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
        reinterpret_cast<uint4*>(o)[j] = q;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (col + j < p.N) o[j] = __float2bfloat16_rn(x[j]);
    }
  } else {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
    if (col + 32 <= p.N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(o)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else {
      for (int j = 0; j < 32; ++j)
        if (col + j < p.N) o[j] = x[j];
    }
  }
}

template <int BN, int MT, bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tap_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapB, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int b_region_slots = p.b_resident ? p.k_chunks : p.b_slots;

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.a_slots * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + b_region_slots * p.b_slot_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxASlots;
  uint64_t* b_full = a_empty + kMaxASlots;     // b_full[0] doubles as the "all weights landed" barrier
  uint64_t* b_empty = b_full + kMaxBSlots;
  uint64_t* acc_full = b_empty + kMaxBSlots;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_slots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kMaxBSlots; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
    mbar_fence_init();
  }
  const uint32_t tmem_cols = static_cast<uint32_t>(BN * MT * p.acc_stages);
  if (warp == 2) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    bool b_loaded = false;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile % p.m_tiles) * 128 * MT;
      const int n0 = (tile / p.m_tiles) * BN;
      if (p.b_resident && !b_loaded) {   // n_tiles == 1 in this mode: load every weight chunk once
        if (elect_one()) {
          mbar_expect_tx(&b_full[0], static_cast<uint32_t>(p.k_chunks * p.b_slot_bytes));
          for (int q = 0; q < p.k_chunks; ++q)
            tma_load_2d(smem_b + q * p.b_slot_bytes, &mapB, &b_full[0], q * p.chunk_elems, n0);
        }
        __syncwarp();
        b_loaded = true;
      }
      for (int g = 0; g < p.ngroups; ++g) {
        const CUtensorMap* mapA = p.g_map[g] ? &mapA1 : &mapA0;
        const int row0 = m0 + p.g_delta[g] - p.halo;
        for (int c = 0; c < p.g_chunks[g]; ++c) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&a_full[as], static_cast<uint32_t>(p.a_slot_bytes));
            uint8_t* dst = smem_a + as * p.a_slot_bytes;
            const int x = p.g_acol[g] + c * p.chunk_elems;
            for (int b = 0; b < p.nboxes; ++b)
              tma_load_2d(dst + b * p.box_rows * kChunkBytes, mapA, &a_full[as], x, row0 + b * p.box_rows);
          }
          __syncwarp();
          if (++as == p.a_slots) { as = 0; aph ^= 1u; }
          if (!p.b_resident) {
            for (int t = 0; t < p.g_ntaps[g]; ++t) {
              mbar_wait(&b_empty[bs], bph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(&b_full[bs], static_cast<uint32_t>(p.b_slot_bytes));
                const int kcol = p.g_kbase[g] + (t * p.g_chunks[g] + c) * p.chunk_elems;
                tma_load_2d(smem_b + bs * p.b_slot_bytes, &mapB, &b_full[bs], kcol, n0);
              }
              __syncwarp();
              if (++bs == p.b_slots) { bs = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops convergently, one elected lane issues) =====
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, accph = 0;
    const uint32_t a_base0 = smem_u32(smem_a);
    const uint32_t b_base0 = smem_u32(smem_b);
    // descriptor high word is constant: SBO = 1024 B, version 1, SWIZZLE_128B (see umma_desc_sw128)
    constexpr uint32_t kDescHi = static_cast<uint32_t>(umma_desc_sw128_hi());
    bool b_ready = false;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], accph ^ 1u);   // epilogue has drained this accumulator stage
      if (p.b_resident && !b_ready) {
        mbar_wait(&b_full[0], 0);
        b_ready = true;
      }
      tc_fence_after();
      const uint32_t d_tile = tmem_base + acc * (BN * MT);
      uint32_t fresh = 1;  // first MMA of each accumulator overwrites, the rest accumulate
      for (int g = 0; g < p.ngroups; ++g) {
        const int nchunks = p.g_chunks[g], ntaps = p.g_ntaps[g], tap0 = p.g_tap0[g];
        const int q0 = p.g_kbase[g] / p.chunk_elems;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_win = a_base0 + as * p.a_slot_bytes;
          for (int t = 0; t < ntaps; ++t) {
            uint32_t b_tile;
            if (p.b_resident) {
              b_tile = b_base0 + (q0 + t * nchunks + c) * p.b_slot_bytes;
            } else {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              b_tile = b_base0 + bs * p.b_slot_bytes;
            }
            const uint32_t a_lo = ((a_win + static_cast<uint32_t>(p.tap_rel[tap0 + t]) * kChunkBytes) & 0x3FFFFu) >> 4;
            const uint32_t b_lo = (b_tile & 0x3FFFFu) >> 4;
            if (elect_one()) {
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 4 x (K = 32 bytes) per 128-byte chunk; 32 B >> 4 = 2
                  const uint64_t ad = (static_cast<uint64_t>(kDescHi) << 32) | (a_lo + sub * (128 * kChunkBytes / 16) + 2 * k);
                  const uint64_t bd = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + 2 * k);
                  const uint32_t accum = (k == 0) ? (fresh ^ 1u) : 1u;
                  if (TF32) umma_tf32(d_tile + sub * BN, ad, bd, p.idesc, accum);
                  else      umma_f16(d_tile + sub * BN, ad, bd, p.idesc, accum);
                }
              }
            }
            __syncwarp();
            fresh = 0;
            if (!p.b_resident) {
              if (elect_one()) umma_commit(&b_empty[bs]);   // weight slot is free once these MMAs retire
              __syncwarp();
              if (++bs == p.b_slots) { bs = 0; bph ^= 1u; }
            }
          }
          if (elect_one()) umma_commit(&a_empty[as]);       // window slot is free once all its taps retire
          __syncwarp();
          if (++as == p.a_slots) { as = 0; aph ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(&acc_full[acc]);
      __syncwarp();
      if (++acc == p.acc_stages) { acc = 0; accph ^= 1u; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = ew >> 2;           // which half of the tile's columns
    constexpr int kCols = BN / 2;       // columns per warp per sub-tile
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile % p.m_tiles) * 128 * MT;
      const int n0 = (tile / p.m_tiles) * BN;
      const int row_first = m0 + quad * 32 + lane;
      ResRegs rr;
      res_prefetch(rr, p, row_first, n0 + half * kCols, row_first < p.M);
      mbar_wait(&acc_full[acc], accph);
      tc_fence_after();
      for (int sub = 0; sub < MT; ++sub) {
        const int row = m0 + sub * 128 + quad * 32 + lane;
        const bool row_ok = row < p.M;
        bool pix_ok = true;
        if (p.mask_en) {
          const int rem = row % p.mRPI;
          pix_ok = (rem / p.mP) < p.mH && (rem % p.mP) < p.mW;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * (BN * MT) + sub * BN +
                               half * kCols;
#pragma unroll 1
        for (int c0 = 0; c0 < kCols; c0 += 32) {
          const int col = n0 + half * kCols + c0;
          uint32_t v[32];
          __syncwarp();                               // tcgen05.ld is warp-collective (.sync.aligned)
          tmem_ld32(taddr + c0, v);
          // prefetch the residual of the next chunk while the TMEM load is in flight
          ResRegs rn;
          {
            int nrow = row, ncol = col + 32;
            if (c0 + 32 >= kCols) { nrow = row + 128; ncol = n0 + half * kCols; }
            const bool more = (c0 + 32 < kCols) || (sub + 1 < MT);
            res_prefetch(rn, p, nrow, ncol, more && nrow < p.M && ncol < p.N);
          }
          tmem_ld_wait();
          if (row_ok && col < p.N) {
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            if (p.bias) {
              if (col + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
                  x[4 * j] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col + j < p.N) x[j] += __ldg(p.bias + col + j);
              }
            }
            res_apply(x, rr, p, row, col);
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
            }
            if (!pix_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = 0.f;   // keep the shared zero padding of the grid intact
            }
            if (p.round_tf32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = round_tf32_rna(x[j]);
            }
            store_row32(x, p, row, col);
          }
          rr = rn;
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == p.acc_stages) { acc = 0; accph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D K-major tensor map: dim0 = cols (contiguous), dim1 = rows with stride ld elements;
// box = {128 bytes, box_rows}, SWIZZLE_128B, out-of-range elements read as zero.
int encode_2d(CUtensorMap* map, bool tf32, uint64_t base, int rows, int cols, int ld, int box_rows,
              const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  VQA_REQUIRE(fn != nullptr, VQA_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const int esz = tf32 ? 4 : 2;
  VQA_REQUIRE((base & 15) == 0, VQA_E_ALIGN, std::string(what) + ": tensor base must be 16-byte aligned");
  VQA_REQUIRE((static_cast<long long>(ld) * esz) % 16 == 0, VQA_E_ALIGN,
              std::string(what) + ": row stride must be a multiple of 16 bytes");
  VQA_REQUIRE(box_rows >= 1 && box_rows <= 256, VQA_E_INVALID, std::string(what) + ": box rows out of range");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kChunkBytes / esz), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  reinterpret_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vqa_set_error(std::string(what) + ": cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
                  " (rows=" + std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld) +
                  " box_rows=" + std::to_string(box_rows) + ")");
    return VQA_E_CUDA;
  }
  return VQA_OK;
}

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor bit layout): c_format f32 [4,6)=1,
// a/b format [7,10)/[10,13) (1 = bf16, 2 = tf32), K-major A and B, N>>3 at [17,23), M>>4 at [24,29).
uint32_t make_idesc(bool tf32, int n) {
  const uint32_t fmt = tf32 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

int num_sms(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}

}  // namespace

typedef void (*GemmKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);

static GemmKernelFn pick_kernel(int bn, int mt, bool tf32) {
#define VQA_PICK(BN_, MT_)                                                                    \
  if (bn == BN_ && mt == MT_) return tf32 ? static_cast<GemmKernelFn>(&gemm_tap_kernel<BN_, MT_, true>) \
                                          : static_cast<GemmKernelFn>(&gemm_tap_kernel<BN_, MT_, false>);
  VQA_PICK(64, 1) VQA_PICK(64, 2) VQA_PICK(128, 1) VQA_PICK(128, 2) VQA_PICK(256, 1) VQA_PICK(256, 2)
#undef VQA_PICK
  return nullptr;
}

struct GemmLaunch {
  GemmKernelFn fn;
  CUtensorMap mapA0, mapA1, mapB;
  GemmParams prm;
  dim3 grid;
  int bn;
  size_t smem;
  uint64_t out_raw, res_raw;
};

int gemm_launch_bytes() { return static_cast<int>(sizeof(GemmLaunch)); }

int gemm_prepare(const VqaOp& op, void* storage, int device) {
  GemmLaunch* L = new (storage) GemmLaunch();
  GemmParams& p = L->prm;
  const int32_t* I = op.i;
  const bool tf32 = I[GEMM_I_dtype] == 1;
  const int bn = I[GEMM_I_BN];
  VQA_REQUIRE(bn == 64 || bn == 128 || bn == 256, VQA_E_INVALID, "gemm: BN must be 64, 128 or 256");
  p.M = I[GEMM_I_M];
  p.N = I[GEMM_I_N];
  p.MT = I[GEMM_I_MT];
  p.halo = I[GEMM_I_halo];
  VQA_REQUIRE(p.MT == 1 || p.MT == 2, VQA_E_INVALID, "gemm: p.MT must be 1 or 2");
  VQA_REQUIRE(p.MT * bn <= 512, VQA_E_INVALID, "gemm: accumulators exceed 512 TMEM columns");
  VQA_REQUIRE(p.M > 0 && p.N > 0 && I[GEMM_I_Npad] % bn == 0 && I[GEMM_I_Npad] >= p.N, VQA_E_INVALID,
              "gemm: bad M/N/Npad");
  p.ngroups = I[GEMM_I_ngroups];
  VQA_REQUIRE(p.ngroups >= 1 && p.ngroups <= VQA_MAX_GROUPS, VQA_E_INVALID, "gemm: bad group count");
  VQA_REQUIRE(I[GEMM_I_ntaps] >= 1 && I[GEMM_I_ntaps] <= VQA_MAX_TAPS, VQA_E_INVALID, "gemm: bad tap count");
  p.chunk_elems = tf32 ? 32 : 64;
  p.is_tf32 = tf32 ? 1 : 0;
  p.idesc = make_idesc(tf32, bn);
  VQA_REQUIRE(I[GEMM_I_Ktot] % p.chunk_elems == 0, VQA_E_INVALID, "gemm: Ktot must be a multiple of the K chunk");
  p.k_chunks = I[GEMM_I_Ktot] / p.chunk_elems;
  bool lockstep = p.halo == 0;
  long long kcover = 0;
  bool uses_a1 = false;
  for (int g = 0; g < p.ngroups; ++g) {
    p.g_map[g] = I[GEMM_I_g_map0 + g];
    p.g_delta[g] = I[GEMM_I_g_delta0 + g];
    p.g_acol[g] = I[GEMM_I_g_acol0 + g];
    p.g_chunks[g] = I[GEMM_I_g_chunks0 + g];
    p.g_ntaps[g] = I[GEMM_I_g_ntaps0 + g];
    p.g_kbase[g] = I[GEMM_I_g_kbase0 + g];
    p.g_tap0[g] = I[GEMM_I_g_tap00 + g];
    VQA_REQUIRE(p.g_chunks[g] >= 1 && p.g_ntaps[g] >= 1 && p.g_tap0[g] + p.g_ntaps[g] <= VQA_MAX_TAPS,
                VQA_E_INVALID, "gemm: bad group");
    VQA_REQUIRE(p.g_kbase[g] % p.chunk_elems == 0, VQA_E_INVALID, "gemm: group K base must be chunk aligned");
    if (p.g_ntaps[g] != 1) lockstep = false;
    uses_a1 |= p.g_map[g] != 0;
    kcover += static_cast<long long>(p.g_ntaps[g]) * p.g_chunks[g] * p.chunk_elems;
  }
  VQA_REQUIRE(kcover == I[GEMM_I_Ktot], VQA_E_INVALID, "gemm: groups do not cover Ktot");
  for (int t = 0; t < VQA_MAX_TAPS; ++t) {
    p.tap_rel[t] = I[GEMM_I_tap_rel0 + t];
    VQA_REQUIRE(t >= I[GEMM_I_ntaps] || (p.tap_rel[t] >= 0 && p.tap_rel[t] <= 2 * p.halo), VQA_E_INVALID,
                "gemm: tap offset outside the window");
  }
  // window geometry
  int win = (128 * p.MT + 2 * p.halo + 7) / 8 * 8;
  p.nboxes = 1;
  p.box_rows = win;
  if (win > 256) {
    p.nboxes = 2;
    p.box_rows = ((win + 1) / 2 + 7) / 8 * 8;
    win = 2 * p.box_rows;
  }
  VQA_REQUIRE(p.box_rows <= 256, VQA_E_INVALID, "gemm: halo too large for a 2-box window");
  p.a_slot_bytes = win * kChunkBytes;
  p.b_slot_bytes = bn * kChunkBytes;
  p.m_tiles = (p.M + 128 * p.MT - 1) / (128 * p.MT);
  p.n_tiles = (p.N + bn - 1) / bn;        // only N tiles that contain real columns run
  p.acc_stages = (2 * p.MT * bn <= 512) ? 2 : 1;

  // shared-memory plan: one CTA per SM (persistent), ~210 KB of rings
  int budget = I[GEMM_I_smem_budget];
  if (budget <= 0) budget = 208 * 1024;
  const long long b_all = static_cast<long long>(p.k_chunks) * p.b_slot_bytes;
  p.b_resident = (p.n_tiles == 1 && b_all <= 96 * 1024 && b_all + 2LL * p.a_slot_bytes <= budget) ? 1 : 0;
  if (p.b_resident) {
    int s = static_cast<int>((budget - b_all) / p.a_slot_bytes);
    p.a_slots = s < 2 ? 2 : (s > kMaxASlots ? kMaxASlots : s);
    p.b_slots = 1;
  } else if (lockstep) {
    int s = budget / (p.a_slot_bytes + p.b_slot_bytes);
    s = s < 2 ? 2 : (s > kMaxASlots ? kMaxASlots : s);
    p.a_slots = p.b_slots = s;
  } else {
    // window mode: 3 windows in flight when they are cheap, the rest of the budget to the weight ring
    p.a_slots = (3 * p.a_slot_bytes + 4 * p.b_slot_bytes <= budget) ? 3 : 2;
    int s = (budget - p.a_slots * p.a_slot_bytes) / p.b_slot_bytes;
    p.b_slots = s < 2 ? 2 : (s > kMaxBSlots ? kMaxBSlots : s);
  }
  const long long b_bytes = p.b_resident ? b_all : static_cast<long long>(p.b_slots) * p.b_slot_bytes;
  L->smem = 1024 + static_cast<size_t>(p.a_slots) * p.a_slot_bytes + static_cast<size_t>(b_bytes) +
            8 * (2 * kMaxASlots + 2 * kMaxBSlots + 4) + 16;
  VQA_REQUIRE(L->smem <= 227 * 1024, VQA_E_INVALID, "gemm: shared memory budget exceeded");

  // tensor maps (only for non-external operands: A and W always live in the arenas)
  VQA_REQUIRE(!(op.p[GEMM_P_a0] & VQA_EXT_TAG) && !(op.p[GEMM_P_b] & VQA_EXT_TAG) && !(op.p[GEMM_P_a1] & VQA_EXT_TAG),
              VQA_E_INVALID, "gemm: A/B operands must be arena buffers");
  int rc = encode_2d(&L->mapA0, tf32, op.p[GEMM_P_a0], I[GEMM_I_a0_rows], I[GEMM_I_a0_cols], I[GEMM_I_a0_ld],
                     p.box_rows, "gemm A0");
  if (rc) return rc;
  if (uses_a1) {
    VQA_REQUIRE(op.p[GEMM_P_a1] != 0, VQA_E_INVALID, "gemm: group references A1 but it is null");
    rc = encode_2d(&L->mapA1, tf32, op.p[GEMM_P_a1], I[GEMM_I_a1_rows], I[GEMM_I_a1_cols], I[GEMM_I_a1_ld],
                   p.box_rows, "gemm A1");
    if (rc) return rc;
  } else {
    L->mapA1 = L->mapA0;
  }
  rc = encode_2d(&L->mapB, tf32, op.p[GEMM_P_b], I[GEMM_I_Npad], I[GEMM_I_Ktot], I[GEMM_I_Ktot], bn, "gemm B");
  if (rc) return rc;

  p.ldo = I[GEMM_I_ldo];
  p.ldr = I[GEMM_I_ldr];
  p.out_dtype = I[GEMM_I_out_dtype];
  p.res_dtype = I[GEMM_I_res_dtype];
  p.relu = I[GEMM_I_relu];
  p.round_tf32 = I[GEMM_I_round_tf32];
  p.mask_en = I[GEMM_I_mask_en];
  p.mP = I[GEMM_I_mP] > 0 ? I[GEMM_I_mP] : 1;
  p.mRPI = I[GEMM_I_mRPI] > 0 ? I[GEMM_I_mRPI] : 1;
  p.mH = I[GEMM_I_mH];
  p.mW = I[GEMM_I_mW];
  p.bias = reinterpret_cast<const float*>(op.p[GEMM_P_bias]);
  VQA_REQUIRE(p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, VQA_E_ALIGN,
              "gemm: bias must be 16-byte aligned");
  L->out_raw = op.p[GEMM_P_out];
  L->res_raw = op.p[GEMM_P_res];
  VQA_REQUIRE(L->out_raw != 0, VQA_E_INVALID, "gemm: null output");
  L->bn = bn;
  const int tiles = p.m_tiles * p.n_tiles;
  const int sms = num_sms(device);
  L->grid = dim3(tiles < sms ? tiles : sms, 1, 1);

  L->fn = pick_kernel(bn, p.MT, tf32);
  VQA_REQUIRE(L->fn != nullptr, VQA_E_INVALID, "gemm: no kernel instantiation for this BN/MT");
  VQA_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(L->fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
  return VQA_OK;
}

int gemm_run(const void* storage, const uint64_t* ext, int n_ext, cudaStream_t stream) {
  const GemmLaunch* L = reinterpret_cast<const GemmLaunch*>(storage);
  GemmParams p = L->prm;
  p.out = reinterpret_cast<void*>(vqa_resolve(L->out_raw, ext, n_ext));
  p.res = reinterpret_cast<const void*>(vqa_resolve(L->res_raw, ext, n_ext));
  VQA_REQUIRE(p.out != nullptr, VQA_E_INVALID, "gemm: unresolved external output");
  L->fn<<<L->grid, kThreads, L->smem, stream>>>(L->mapA0, L->mapA1, L->mapB, p);
  VQA_LAUNCH_OK("gemm_tap_kernel");
  return VQA_OK;
}

const char* gemm_kernel_name(const void* storage) {
  const GemmLaunch* L = reinterpret_cast<const GemmLaunch*>(storage);
  static thread_local char name[64];
  snprintf(name, sizeof(name), "gemm_tap_kernel<%d,%d,%s>", L->bn, L->prm.MT, L->prm.is_tf32 ? "tf32" : "bf16");
  return name;
}
