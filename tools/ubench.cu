// B200 micro-benchmarks behind the design of the shift-fused convolution kernels:
//   ldtm   tcgen05.ld 32x32b.x32 throughput per SM with 4 / 8 / 16 reader warps
//   shfl   warp-shuffle throughput per SM with 8 warps
//   umma   cycles per SS-mode tcgen05.mma (M = 128 or 256 with cta_group::2) for N = 64 .. 256, K = 16, issued back to
//          back over row-shifted views of one A window like gemm_tap_kernel does, alone and with epilogue warps reading
//          TMEM at the same time
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench tools/ubench.cu
// (test / profiling aid; nothing in the product links it)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../visual-question-answering-vqa-system_b200/csrc/common.cuh"

void vqa_set_error(const std::string& msg) { fprintf(stderr, "%s\n", msg.c_str()); }
void vqa_count_launch() {}
bool vqa_pdl_enabled() { return false; }

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e__ = (x);                                                         \
    if (e__ != cudaSuccess) {                                                      \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------------ ldtm
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tslot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[32], b[32];
    tmem_ld32(base + ((2 * i) & 15) * 32, a);
    tmem_ld32(base + ((2 * i + 1) & 15) * 32, b);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 32; ++k) acc ^= a[k] + b[k];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tslot, 512);
}

// ------------------------------------------------------------------------------------------------ shfl
__global__ void __launch_bounds__(256, 1) shfl_kernel(int iters, long long* out, float* sink) {
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = threadIdx.x * 0.5f + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __shfl_down_sync(0xffffffffu, v[k], 1) + 1.0f;
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) s += v[k];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 123.456f) sink[0] = s;
}

// ------------------------------------------------------------------------------------------------ umma
struct UmmaCfg {
  int N, MT, ntaps, tap_stride_rows, iters, ld_warps, pair;
};

template <bool PAIR>
__global__ void __launch_bounds__(320, 1) umma_kernel(UmmaCfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const int a_bytes = (128 * c.MT + 2 * 64 + 8) * 128;                 // window: tile rows + 64 rows of halo either side
  const int b_slot = (PAIR ? c.N / 2 : c.N) * 128;
  uint8_t* sa = smem;
  uint8_t* sb = smem + (a_bytes + 1023) / 1024 * 1024;
  // plausible operand values (power draw depends on the data): small bf16 numbers from an LCG
  {
    uint32_t s = 0x9e3779b9u * (threadIdx.x + 1) + blockIdx.x;
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    const int words = ((a_bytes + 1023) / 1024 * 1024 + c.ntaps * b_slot) / 4;
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      s = s * 1664525u + 1013904223u;
      w[i] = 0x3c003c00u ^ (s & 0x807f807fu);
    }
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); done = 0; }
  if (warp == 2) {
    if (PAIR) { tmem_alloc_pair(&tslot, 512); tmem_relinquish_pair(); }
    else      { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  }
  fence_proxy_async();
  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  long long cyc = 0, lds = 0;
  if (warp == 1) {
    if ((!PAIR || rank == 0) && elect_one()) {
      const uint64_t hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(c.N >> 3) << 17) |
                             (static_cast<uint32_t>((PAIR ? 256 : 128) >> 4) << 24);
      const uint32_t a_lo = smem_u32(sa) >> 4, b_lo = smem_u32(sb) >> 4;
      const long long t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        for (int t = 0; t < c.ntaps; ++t) {
          const uint32_t arel = static_cast<uint32_t>(t * c.tap_stride_rows) * 8u;   // rows * 128 B >> 4
          for (int sub = 0; sub < c.MT; ++sub) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = hi | (a_lo + arel + sub * 1024 + 2 * k);
              const uint64_t bd = hi | (b_lo + t * (b_slot >> 4) + 2 * k);
              if (PAIR) umma_f16_pair(tbase + sub * c.N, ad, bd, idesc, 1u);
              else      umma_f16(tbase + sub * c.N, ad, bd, idesc, 1u);
            }
          }
        }
      }
      if (PAIR) umma_commit_pair(&bar); else umma_commit(&bar);
      mbar_wait(&bar, 0);
      cyc = clock64() - t0;
      done = 1;
      out[blockIdx.x * 2] = cyc;
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 2 + c.ld_warps) {
    // epilogue-like TMEM readers running while the MMAs execute (they read the accumulator columns being written)
    const uint32_t base = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    if (PAIR && rank != 0) mbar_wait(&bar, 0);   // follower: just wait for the leader's commit
    else {
      while (!done) {
        uint32_t a[32], b[32];
        tmem_ld32(base + ((2 * lds) & 7) * 32, a);
        tmem_ld32(base + ((2 * lds + 1) & 7) * 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= a[k] + b[k];
        ++lds;
      }
    }
    if (acc == 0x12345678u) out[0] = 1;
    if (warp == 2 && lane == 0) out[blockIdx.x * 2 + 1] = lds;
  }
  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tbase, 512); else tmem_dealloc(tbase, 512);
  }
}

static void run_umma(const UmmaCfg& c, long long* d_out, int sms) {
  const int a_bytes = ((128 * c.MT + 2 * 64 + 8) * 128 + 1023) / 1024 * 1024;
  const int b_slot = (c.pair ? c.N / 2 : c.N) * 128;
  const size_t smem = 1024 + a_bytes + static_cast<size_t>(c.ntaps) * b_slot;
  if (smem > 226 * 1024) { printf("umma N=%d skipped (smem)\n", c.N); return; }
  CK(cudaMemset(d_out, 0, sizeof(long long) * 2 * sms));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    if (c.pair) {
      CK(cudaFuncSetAttribute(umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      CK(vqa_launch_cluster(umma_kernel<true>, dim3(sms / 2 * 2), dim3(320), smem, 0, 2, c, d_out));
    } else {
      CK(cudaFuncSetAttribute(umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      CK(vqa_launch_cluster(umma_kernel<false>, dim3(sms), dim3(320), smem, 0, 1, c, d_out));
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(2 * sms);
  CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
  const long long n_mma = static_cast<long long>(c.iters) * c.ntaps * c.MT * 4;
  const double cyc = static_cast<double>(h[0]) / n_mma;
  const double math = (c.pair ? 256.0 : 128.0) * c.N * 16 / (c.pair ? 8192.0 : 4096.0);
  const double lds_bytes = static_cast<double>(h[1]) * 2 * 4096 * c.ld_warps;
  printf("umma %s M=%3d N=%3d MT=%d taps=%d ld_warps=%2d : %7.1f clk/MMA (math floor %5.1f, %4.0f %%)  wall %.3f ms  "
         "clk %.0f MHz  concurrent LDTM %.1f B/clk/SM\n",
         c.pair ? "pair" : "1cta", c.pair ? 256 : 128, c.N, c.MT, c.ntaps, c.ld_warps, cyc, math, 100.0 * math / cyc, ms,
         h[0] / (ms * 1e3), h[0] ? lds_bytes / h[0] : 0.0);
}

int main() {
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d_out;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_out, sizeof(long long) * 4 * sms));
  CK(cudaMalloc(&d_sink, 64));
  std::vector<long long> h(sms);
  for (int warps : {4, 8, 16}) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) {
      ldtm_kernel<<<sms, warps * 32>>>(iters, d_out, d_sink);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    printf("ldtm 32x32b.x32, %2d warps/SM: %.1f B/clk/SM (%lld clk for %d x 2 loads per warp)\n", warps,
           static_cast<double>(warps) * iters * 2 * 4096 / h[0], h[0], iters);
  }
  {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      shfl_kernel<<<sms, 256>>>(iters, d_out, reinterpret_cast<float*>(d_sink));
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    printf("shfl.down + fadd, 8 warps/SM: %.2f warp-shuffles/clk/SM\n", 8.0 * iters * 32 / h[0]);
  }
  const int iters = 300;
  for (int ldw : {0, 8}) {
    for (int pair : {0, 1}) {
      run_umma({64, 2, 9, 7, iters, ldw, pair}, d_out, sms);
      run_umma({128, 2, 9, 7, iters, ldw, pair}, d_out, sms);
      run_umma({192, 1, 3, 57, 3 * iters, ldw, pair}, d_out, sms);
      run_umma({192, 2, 3, 57, 3 * iters, ldw, pair}, d_out, sms);
      run_umma({256, 1, 4, 30, 2 * iters, ldw, pair}, d_out, sms);
      run_umma({128, 2, 8, 14, iters, ldw, pair}, d_out, sms);
    }
  }
  printf("done\n");
  return 0;
}
