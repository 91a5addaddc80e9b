// Probe of tcgen05.shift.down on B200: which rows / columns move, and how long one shift takes in the tensor pipe.
// (test aid behind DESIGN.md section 3.1; nothing in the product links it)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../visual-question-answering-vqa-system_b200/csrc/common.cuh"

void vqa_set_error(const std::string& msg) { fprintf(stderr, "%s\n", msg.c_str()); }
void vqa_count_launch() {}
bool vqa_pdl_enabled() { return false; }

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); exit(1); } } while (0)

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
__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}

// mode 0: semantics (one shift at column `col`, lane base `lane0`); mode 1: timing of n back-to-back shifts
__global__ void __launch_bounds__(128, 1) shift_kernel(int mode, int col, int lane0, int n, uint32_t* out, long long* cyc) {
  __shared__ uint32_t tslot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tslot, 64); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  const uint32_t mine = tb + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = (threadIdx.x << 8) | k;   // value = row * 256 + column
  tmem_st32(mine, v);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) tmem_shift_down(tb + (static_cast<uint32_t>(lane0) << 16) + col);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  tmem_ld32(mine, v);
  tmem_ld_wait();
  if (mode == 0) {
#pragma unroll
    for (int k = 0; k < 32; ++k) out[threadIdx.x * 32 + k] = v[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

int main() {
  uint32_t* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, 128 * 32 * 4));
  CK(cudaMalloc(&d_cyc, 64));
  std::vector<uint32_t> h(128 * 32);
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int col = cfg == 1 ? 8 : 0, lane0 = cfg == 2 ? 32 : 0;
    shift_kernel<<<1, 128>>>(0, col, lane0, 1, d_out, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
    printf("one shift.down at column %d, lane base %d -- rows whose value changed (row: column -> source row):\n", col, lane0);
    int shown = 0;
    for (int r = 0; r < 128; ++r) {
      int first = -1, last = -1, src = -1;
      bool uniform = true;
      for (int k = 0; k < 32; ++k) {
        const uint32_t want = (r << 8) | k;
        if (h[r * 32 + k] != want) {
          if (first < 0) { first = k; src = h[r * 32 + k] >> 8; }
          last = k;
          if (static_cast<int>(h[r * 32 + k] >> 8) != src || static_cast<int>(h[r * 32 + k] & 255) != k) uniform = false;
        }
      }
      if (first >= 0 && (r < 4 || (r % 32) < 2 || (r % 32) > 29 || r > 124) && shown < 40) {
        printf("  row %3d: columns %d..%d <- row %d%s\n", r, first, last, src, uniform ? "" : " (mixed)");
        ++shown;
      }
    }
    int changed = 0;
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) changed += h[r * 32 + k] != static_cast<uint32_t>((r << 8) | k);
    printf("  %d values changed in total\n", changed);
  }
  for (int n : {1, 8, 64, 256}) {
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) {
      shift_kernel<<<1, 128>>>(1, 0, 0, n, d_out, d_cyc);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("%3d shifts + commit + wait: %lld clk (%.1f per shift)\n", n, c, static_cast<double>(c) / n);
  }
  return 0;
}
