// Tap-shifted GEMM on tcgen05 / TMEM fed by TMA  (VQA_OP_GEMM).
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
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM -> registers
// -> bias / residual / ReLU / pad-mask -> global).  Accumulators: MT sub-tiles of 128 x BN fp32 in
// TMEM.  Pipelines: A-window ring and B (weight) ring, each with full/empty mbarriers.
#include "common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kChunkBytes = 128;  // one SWIZZLE_128B row: 64 bf16 or 32 fp32 (tf32)

struct GemmParams {
  int M, N;
  int MT;            // 128-row sub-tiles per CTA (1 or 2)
  int halo;          // window rows before/after the tile
  int box_rows;      // TMA box rows for A
  int nboxes;        // boxes per window (1 or 2)
  int a_slots, b_slots;
  int a_slot_bytes, b_slot_bytes;
  int ngroups;
  int chunk_elems;   // 64 (bf16) / 32 (tf32)
  int is_tf32;
  uint32_t idesc;
  int desc_mode;     // 0: base_offset field 0 (absolute-address swizzle); 1: base_offset=(addr>>7)&7
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

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int desc_mode) {
  uint64_t d = umma_desc_sw128(addr);
  if (desc_mode == 1) d |= static_cast<uint64_t>((addr >> 7) & 7u) << 49;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tap_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapB, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128 * p.MT;
  const int n0 = blockIdx.y * BN;

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.a_slots * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.b_slots * p.b_slot_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_slots;
  uint64_t* b_full = a_empty + p.a_slots;
  uint64_t* b_empty = b_full + p.b_slots;
  uint64_t* acc_full = b_empty + p.b_slots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_slots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_slots; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(BN * p.MT));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int g = 0; g < p.ngroups; ++g) {
        const CUtensorMap* mapA = p.g_map[g] ? &mapA1 : &mapA0;
        const int row0 = m0 + p.g_delta[g] - p.halo;
        for (int c = 0; c < p.g_chunks[g]; ++c) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          mbar_expect_tx(&a_full[as], static_cast<uint32_t>(p.a_slot_bytes));
          uint8_t* dst = smem_a + as * p.a_slot_bytes;
          const int x = p.g_acol[g] + c * p.chunk_elems;
          for (int b = 0; b < p.nboxes; ++b)
            tma_load_2d(dst + b * p.box_rows * kChunkBytes, mapA, &a_full[as], x, row0 + b * p.box_rows);
          if (++as == p.a_slots) { as = 0; aph ^= 1u; }
          for (int t = 0; t < p.g_ntaps[g]; ++t) {
            mbar_wait(&b_empty[bs], bph ^ 1u);
            mbar_expect_tx(&b_full[bs], static_cast<uint32_t>(p.b_slot_bytes));
            const int kcol = p.g_kbase[g] + (t * p.g_chunks[g] + c) * p.chunk_elems;
            tma_load_2d(smem_b + bs * p.b_slot_bytes, &mapB, &b_full[bs], kcol, n0);
            if (++bs == p.b_slots) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      uint32_t first = 1;  // first MMA of each accumulator overwrites, the rest accumulate
      const uint32_t a_base0 = smem_u32(smem_a);
      const uint32_t b_base0 = smem_u32(smem_b);
      for (int g = 0; g < p.ngroups; ++g) {
        for (int c = 0; c < p.g_chunks[g]; ++c) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_win = a_base0 + as * p.a_slot_bytes;
          for (int t = 0; t < p.g_ntaps[g]; ++t) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_tile = b_base0 + bs * p.b_slot_bytes;
            const uint32_t a_tap = a_win + static_cast<uint32_t>(p.tap_rel[p.g_tap0[g] + t]) * kChunkBytes;
            for (int sub = 0; sub < p.MT; ++sub) {
              const uint32_t a_sub = a_tap + sub * 128 * kChunkBytes;
              const uint32_t d_tmem = tmem_base + sub * BN;
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // 4 x (K = 32 bytes) per 128-byte chunk
                const uint64_t ad = make_desc(a_sub + k * 32, p.desc_mode);
                const uint64_t bd = make_desc(b_tile + k * 32, p.desc_mode);
                const uint32_t acc = (first && k == 0) ? 0u : 1u;
                if (p.is_tf32) umma_tf32(d_tmem, ad, bd, p.idesc, acc);
                else           umma_f16(d_tmem, ad, bd, p.idesc, acc);
              }
            }
            first = 0;
            umma_commit(&b_empty[bs]);   // weight slot is free once these MMAs retire
            if (++bs == p.b_slots) { bs = 0; bph ^= 1u; }
          }
          umma_commit(&a_empty[as]);     // window slot is free once all its taps retire
          if (++as == p.a_slots) { as = 0; aph ^= 1u; }
        }
      }
      umma_commit(acc_full);
    }
  } else {
    // ===================== epilogue warps (TMEM lane quadrant = warp % 4) =====================
    const int quad = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool out_bf16 = p.out_dtype == 0;
    for (int sub = 0; sub < p.MT; ++sub) {
      const int row = m0 + sub * 128 + quad * 32 + lane;
      bool row_ok = row < p.M;
      bool pix_ok = true;
      if (p.mask_en) {
        const int rem = row % p.mRPI;
        pix_ok = (rem / p.mP) < p.mH && (rem % p.mP) < p.mW;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + sub * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        const int col = n0 + c0;
        if (col >= p.N) break;                      // warp-uniform
        uint32_t v[16];
        __syncwarp();                               // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (row_ok) {                               // rows past M only take part in the TMEM load
        float x[16];
        const bool full = col + 16 <= p.N;
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] += (full || col + j < p.N) ? __ldg(p.bias + col + j) : 0.f;
        }
        if (p.res) {
          if (p.res_dtype == 0) {
            const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(p.res) + static_cast<size_t>(row) * p.ldr + col;
            if (full && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
              const uint4 q0 = *reinterpret_cast<const uint4*>(r);
              const uint4 q1 = *reinterpret_cast<const uint4*>(r + 8);
              const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                x[2 * j] += __uint_as_float(w[j] << 16);
                x[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
              }
            } else {
              for (int j = 0; j < 16; ++j) if (col + j < p.N) x[j] += __bfloat162float(r[j]);
            }
          } else {
            const float* r = reinterpret_cast<const float*>(p.res) + static_cast<size_t>(row) * p.ldr + col;
            if (full && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 q = *reinterpret_cast<const float4*>(r + 4 * j);
                x[4 * j] += q.x; x[4 * j + 1] += q.y; x[4 * j + 2] += q.z; x[4 * j + 3] += q.w;
              }
            } else {
              for (int j = 0; j < 16; ++j) if (col + j < p.N) x[j] += r[j];
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (!pix_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = 0.f;   // keep the shared zero padding of the grid intact
        }
        if (p.round_tf32) {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = round_tf32_rna(x[j]);
        }
        if (out_bf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
          if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            uint4 q0, q1;
            q0.x = pack_bf16x2(x[0], x[1]);   q0.y = pack_bf16x2(x[2], x[3]);
            q0.z = pack_bf16x2(x[4], x[5]);   q0.w = pack_bf16x2(x[6], x[7]);
            q1.x = pack_bf16x2(x[8], x[9]);   q1.y = pack_bf16x2(x[10], x[11]);
            q1.z = pack_bf16x2(x[12], x[13]); q1.w = pack_bf16x2(x[14], x[15]);
            *reinterpret_cast<uint4*>(o) = q0;
            *reinterpret_cast<uint4*>(o + 8) = q1;
          } else {
            for (int j = 0; j < 16; ++j) if (col + j < p.N) o[j] = __float2bfloat16_rn(x[j]);
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
          if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(o + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          } else {
            for (int j = 0; j < 16; ++j) if (col + j < p.N) o[j] = x[j];
          }
        }
        }  // row_ok
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, static_cast<uint32_t>(BN * p.MT));
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

}  // namespace

struct GemmLaunch {
  CUtensorMap mapA0, mapA1, mapB;
  GemmParams prm;
  dim3 grid;
  int bn;
  size_t smem;
  bool ext_out, ext_res;
  uint64_t out_raw, res_raw;
};

int gemm_launch_bytes() { return static_cast<int>(sizeof(GemmLaunch)); }

int gemm_prepare(const VqaOp& op, void* storage) {
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
  VQA_REQUIRE(p.MT == 1 || p.MT == 2, VQA_E_INVALID, "gemm: MT must be 1 or 2");
  VQA_REQUIRE(p.MT * bn <= 512, VQA_E_INVALID, "gemm: accumulators exceed 512 TMEM columns");
  VQA_REQUIRE(p.M > 0 && p.N > 0 && I[GEMM_I_Npad] % bn == 0 && I[GEMM_I_Npad] >= p.N, VQA_E_INVALID,
              "gemm: bad M/N/Npad");
  p.ngroups = I[GEMM_I_ngroups];
  VQA_REQUIRE(p.ngroups >= 1 && p.ngroups <= VQA_MAX_GROUPS, VQA_E_INVALID, "gemm: bad group count");
  VQA_REQUIRE(I[GEMM_I_ntaps] >= 1 && I[GEMM_I_ntaps] <= VQA_MAX_TAPS, VQA_E_INVALID, "gemm: bad tap count");
  p.chunk_elems = tf32 ? 32 : 64;
  p.is_tf32 = tf32 ? 1 : 0;
  p.idesc = make_idesc(tf32, bn);
  p.desc_mode = I[GEMM_I_desc_mode];
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
  // pipeline depth from the shared-memory budget
  int n_iters = 0;
  for (int g = 0; g < p.ngroups; ++g) n_iters += p.g_chunks[g];
  int budget = I[GEMM_I_smem_budget];
  if (budget <= 0) budget = (lockstep && n_iters <= 16 && p.MT * bn <= 256) ? 100 * 1024 : 200 * 1024;
  if (lockstep) {
    int s = budget / (p.a_slot_bytes + p.b_slot_bytes);
    s = s < 2 ? 2 : (s > 8 ? 8 : s);
    p.a_slots = p.b_slots = s;
  } else {
    p.a_slots = 2;
    int s = (budget - 2 * p.a_slot_bytes) / p.b_slot_bytes;
    p.b_slots = s < 2 ? 2 : (s > 10 ? 10 : s);
  }
  L->smem = 1024 + static_cast<size_t>(p.a_slots) * p.a_slot_bytes + static_cast<size_t>(p.b_slots) * p.b_slot_bytes +
            8 * (2 * p.a_slots + 2 * p.b_slots + 1) + 16;
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
  L->out_raw = op.p[GEMM_P_out];
  L->res_raw = op.p[GEMM_P_res];
  VQA_REQUIRE(L->out_raw != 0, VQA_E_INVALID, "gemm: null output");
  L->bn = bn;
  L->grid = dim3((p.M + 128 * p.MT - 1) / (128 * p.MT), I[GEMM_I_Npad] / bn, 1);
  // only N tiles that contain real columns need to run
  L->grid.y = (p.N + bn - 1) / bn;

  auto set_attr = [&](const void* fn) -> int {
    VQA_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return VQA_OK;
  };
  if (bn == 64) rc = set_attr(reinterpret_cast<const void*>(&gemm_tap_kernel<64>));
  else if (bn == 128) rc = set_attr(reinterpret_cast<const void*>(&gemm_tap_kernel<128>));
  else rc = set_attr(reinterpret_cast<const void*>(&gemm_tap_kernel<256>));
  return rc;
}

int gemm_run(const void* storage, const uint64_t* ext, int n_ext, cudaStream_t stream) {
  const GemmLaunch* L = reinterpret_cast<const GemmLaunch*>(storage);
  GemmParams p = L->prm;
  p.out = reinterpret_cast<void*>(vqa_resolve(L->out_raw, ext, n_ext));
  p.res = reinterpret_cast<const void*>(vqa_resolve(L->res_raw, ext, n_ext));
  VQA_REQUIRE(p.out != nullptr, VQA_E_INVALID, "gemm: unresolved external output");
  if (L->bn == 64)
    gemm_tap_kernel<64><<<L->grid, kThreads, L->smem, stream>>>(L->mapA0, L->mapA1, L->mapB, p);
  else if (L->bn == 128)
    gemm_tap_kernel<128><<<L->grid, kThreads, L->smem, stream>>>(L->mapA0, L->mapA1, L->mapB, p);
  else
    gemm_tap_kernel<256><<<L->grid, kThreads, L->smem, stream>>>(L->mapA0, L->mapA1, L->mapB, p);
  VQA_LAUNCH_OK("gemm_tap_kernel");
  return VQA_OK;
}

const char* gemm_kernel_name(const void* storage) {
  const GemmLaunch* L = reinterpret_cast<const GemmLaunch*>(storage);
  return L->bn == 64 ? "gemm_tap_kernel<64>" : (L->bn == 128 ? "gemm_tap_kernel<128>" : "gemm_tap_kernel<256>");
}
