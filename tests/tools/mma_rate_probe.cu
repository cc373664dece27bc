// Microbenchmark (run on a B200): how many SM cycles does one tcgen05.mma take when nothing else is going on?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I artstyletransfer_b200/csrc \
//        tests/tools/mma_rate_probe.cu -o tests/tools/build/mma_rate_probe
// One CTA per SM (or one CTA pair per TPC); the operands sit in shared memory (never reloaded), one thread issues
// `n_mma` MMAs back to back, commits, and waits; clock64 around the whole thing.  Variants: kind::tf32 (K = 8) and
// kind::f16 with bf16 inputs (K = 16), M = 128 (cta_group::1) or 256 (cta_group::2), N = 128 / 256, K-major SW128
// operands, optionally with `extra` warps hammering shared memory with LDS/STS (the converter + epilogue traffic of
// the Gram kernels) to see how much tensor throughput the shared-memory pipe takes away.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100_ptx.cuh"

using namespace ast::ptx;

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Args {
  int kind;      // 0 tf32, 1 bf16
  int N;         // 128 or 256
  int n_mma;     // MMAs issued back to back
  int extra;     // warps doing LDS.128 + STS.128 over a 64 KB region meanwhile
  int distinct;  // 1: every MMA reads different operand bytes (walks a 96 KB window), 0: the same 12 KB
  long long* cycles;   // per CTA
};

template <int PAIR>
__global__ void __launch_bounds__(32 * 10, 1) probe(Args A) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
    stop = 0;
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2cta(smem_u32(&tmem_slot), 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && lane == 0 && rank == 0) {
    const int M = PAIR ? 256 : 128;
    const uint32_t idesc = A.kind == 0 ? umma_idesc_tf32(M, A.N, 0, 0) : idesc_bf16(M, A.N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < A.n_mma; ++i) {
      // K-major SW128 operands: a stage of 4 K-steps (32 B apart) x 4 stages 16 KB (A) / 32 KB (B) apart
      const int ks = i & 3, st = A.distinct ? ((i >> 2) & 1) : 0;
      const uint64_t ad = umma_desc_sw128(a0 + st * 16384 + ks * 32, 16, 1024);
      const uint64_t bd = umma_desc_sw128(b0 + st * 32768 + ks * 32, 16, 1024);
      const uint32_t d = tmem + ((i & 1) ? 256u : 0u);
      if (A.kind == 0) { if (PAIR) umma_tf32_2cta(d, ad, bd, idesc, 1u); else umma_tf32(d, ad, bd, idesc, 1u); }
      else { if (PAIR) umma_bf16_2cta(d, ad, bd, idesc, 1u); else umma_bf16(d, ad, bd, idesc, 1u); }
    }
    if (PAIR) umma_commit_2cta(smem_u32(&bar), 1); else umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    A.cycles[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (warp >= 2 && warp < 2 + A.extra && !(PAIR && rank != 0 && false)) {
    // shared-memory traffic of the other roles: every thread streams 16-byte loads + stores over a private 64 KB window
    const uint32_t base = smem_u32(smem + 96 * 1024);
    uint32_t off = (uint32_t)(threadIdx.x - 64) * 16u;
    uint32_t acc = 0;
    while (!stop || (PAIR && rank != 0 && acc < 4000000u)) {
      uint32_t a, b, c, d;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + off));
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + off), "r"(a + 1), "r"(b), "r"(c), "r"(d) : "memory");
      off = (off + (uint32_t)A.extra * 32u * 16u) & 0xffffu;
      acc += 1;
      if (PAIR && rank != 0 && acc >= 200000u) break;
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { if (PAIR) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 1024);
  const int smem = 170 * 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<long long> h(1024);
  printf("{\"sms\": %d, \"clock_khz\": %d}\n", sms, khz);
  for (int pair = 0; pair <= 1; ++pair)
    for (int kind = 0; kind <= 1; ++kind)
      for (int N : {128, 256})
        for (int distinct : {0, 1})
          for (int extra : {0, 4, 8}) {
            Args a{kind, N, 4096, extra, distinct, d_cycles};
            const int grid = pair ? (sms / 2) * 2 : sms;
            for (int rep = 0; rep < 2; ++rep) {
              cudaMemset(d_cycles, 0, sizeof(long long) * 1024);
              if (pair) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, probe<1>, a);
              } else {
                probe<0><<<grid, 320, smem>>>(a);
              }
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("{\"error\": \"%s\", \"pair\": %d, \"kind\": %d, \"N\": %d}\n", cudaGetErrorString(e), pair, kind, N); return 1; }
            }
            cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 1024, cudaMemcpyDeviceToHost);
            long long mx = 0, mn = 1LL << 60; int cnt = 0;
            for (int i = 0; i < grid; ++i) if (h[i] > 0) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; ++cnt; }
            const int M = pair ? 256 : 128, K = kind == 0 ? 8 : 16;
            const double cyc = (double)mx / a.n_mma;
            printf("{\"cta_group\": %d, \"kind\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"distinct_operands\": %d, \"smem_traffic_warps\": %d, "
                   "\"cycles_per_mma_max\": %.1f, \"cycles_per_mma_min\": %.1f, \"issuers\": %d, \"flop_per_cycle_per_sm\": %.0f}\n",
                   pair + 1, kind == 0 ? "tf32" : "bf16", M, N, K, distinct, extra, cyc, (double)mn / a.n_mma, cnt,
                   2.0 * M * N * K / cyc / (pair ? 2 : 1));
            fflush(stdout);
          }
  return 0;
}
