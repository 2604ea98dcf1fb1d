// Practical ceiling for the gamete kernel's access pattern on this GPU: groups of 8 lanes read two
// random 256-byte rows and write one 256-byte row, nothing else.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/row_gather_probe tools/row_gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

template <int ROWQ>   // 128-bit units per homologue (row = 2 * ROWQ units)
__global__ void __launch_bounds__(256) probe(const uint4* G, uint4* out, const int* s0, const int* s1, const int* c, int B) {
  const int lane = threadIdx.x & 7;
  const int ngroups = gridDim.x * blockDim.x / 8;
  for (int o = (blockIdx.x * blockDim.x + threadIdx.x) / 8; o < B; o += ngroups) {
    const uint4* P0 = G + (size_t)s0[o] * 2 * ROWQ;
    const uint4* P1 = G + (size_t)s1[o] * 2 * ROWQ;
    uint4* C = out + (size_t)c[o] * 2 * ROWQ;
    for (int q = lane; q < ROWQ; q += 8) {
      const uint4 a0 = ld_stream(P0 + q), a1 = ld_stream(P0 + ROWQ + q);
      const uint4 b0 = ld_stream(P1 + q), b1 = ld_stream(P1 + ROWQ + q);
      st_stream(C + q, make_uint4(a0.x ^ a1.x, a0.y ^ a1.y, a0.z ^ a1.z, a0.w ^ a1.w));
      st_stream(C + ROWQ + q, make_uint4(b0.x ^ b1.x, b0.y ^ b1.y, b0.z ^ b1.z, b0.w ^ b1.w));
    }
  }
}

template <int ROWQ>
void run(int rows, int B) {
  const size_t row_units = 2 * ROWQ;
  uint4 *G, *out;
  cudaMalloc(&G, (size_t)rows * row_units * 16);
  cudaMalloc(&out, (size_t)rows * row_units * 16);
  cudaMemset(G, 1, (size_t)rows * row_units * 16);
  std::vector<int> h0(B), h1(B), hc(B);
  uint64_t x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (int)(x % (uint64_t)rows); };
  for (int i = 0; i < B; ++i) { h0[i] = rnd(); h1[i] = rnd(); hc[i] = rnd(); }
  int *s0, *s1, *c;
  cudaMalloc(&s0, B * 4); cudaMalloc(&s1, B * 4); cudaMalloc(&c, B * 4);
  cudaMemcpy(s0, h0.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(s1, h1.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(c, hc.data(), B * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int per_sm : {4, 8, 16}) {
    probe<ROWQ><<<148 * per_sm, 256>>>(G, out, s0, s1, c, B);
    cudaEventRecord(e0);
    for (int k = 0; k < 5; ++k) probe<ROWQ><<<148 * per_sm, 256>>>(G, out, s0, s1, c, B);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double bytes = (double)B * (3.0 * row_units * 16 + 12);
    printf("row %4zu B  rows %d  births %d  grid %2d/SM: %.1f us  %.0f GB/s\n", row_units * 16, rows, B, per_sm,
           ms * 1e3, bytes / (ms * 1e-3) / 1e9);
  }
  cudaFree(G); cudaFree(out); cudaFree(s0); cudaFree(s1); cudaFree(c);
}

int main() {
  run<8>(12000000, 2000000);     // c4: 256-byte rows
  run<80>(600000, 100000);       // c5: 2560-byte rows
  run<1>(1300000, 210000);       // c2: 32-byte rows
  return 0;
}
