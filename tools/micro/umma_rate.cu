// Micro-benchmark: cycles per tcgen05.mma (M=128, K=32 bytes) as a function of N, operand kind and the number of
// independent TMEM accumulators the issue loop rotates over.  One CTA per SM, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int KIND>   // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void k(int N, int nacc, int reps, uint32_t sbo, int a_off, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t a = smem_u32(smem) + a_off, b = smem_u32(smem) + 32768;
    // descriptors and accumulator addresses are loop-invariant registers: the loop measures the tensor core (operand fetch from
    // shared memory + math), not the issuing thread (with the descriptors built per iteration every case measured 142 cycles)
    uint64_t da[4], db[4];
    for (int i = 0; i < 4; ++i) { da[i] = desc_sw128(a + i * 32, sbo); db[i] = desc_sw128(b + i * 32, 1024); }
    uint32_t dacc[4];
    for (int i = 0; i < 4; ++i) dacc[i] = tm + (uint32_t)((i % nacc) * N);
    for (int i = 0; i < 4; ++i) mma<KIND>(dacc[i], da[i], db[i], idesc, 0);
    long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mma<KIND>(dacc[i], da[i], db[i], idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

// Conv-like pattern: 36 MMAs per "tile" (9 taps x 4 k-steps of a 32-channel chunk) with the operand addresses of the persistent
// halo kernel (A: tap row/column offsets into a 10 x 18 pixel halo tile of 128-byte rows, SBO = 1280; B: 36 distinct 32-byte
// slices of 9 resident 4 KB panels), accumulating into one TMEM accumulator per tile, 4 accumulators in rotation.
// bg: 0 none, 1 eight other warps stream 128-bit shared-memory LOADS (the epilogue's bias / slope tables), 2 they stream 128-bit
// shared-memory STORES into an unrelated region (stand-in for TMA writes of the next halo tiles).
// KSTEPS: K-steps issued per tap (4 = a full 128-byte chunk; 2 = the 32-channel fp16 layer, half of every 128-byte row is padding)
template <int KIND, int KSTEPS>
__global__ void kconv(int N, int tiles, int bg, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < (128 * 1024) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 32768;
    uint64_t da[36], db[36];
    for (int tap = 0; tap < 9; ++tap)
      for (int k = 0; k < 4; ++k) {
        da[tap * 4 + k] = desc_sw128(a + ((tap / 3) * 10 + (tap % 3)) * 128 + k * 32, 1280);
        db[tap * 4 + k] = desc_sw128(b + tap * 4096 + k * 32, 1024);
      }
    long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tm + (uint32_t)((t & 3) * N);
#pragma unroll
      for (int i = 0; i < 36; ++i)
        if ((i & 3) < KSTEPS) mma<KIND>(d, da[i], db[i], idesc, i != 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x >= 64 && bg) {
    // background traffic on [80 KB, 128 KB): away from the operands
    uint4 *reg = (uint4 *)(smem + 80 * 1024);
    uint4 acc = make_uint4(0, 0, 0, 0);
    int i = threadIdx.x - 64;
    while (!stop) {
      if (bg == 1) {
        uint4 v = reg[i & 2047];
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
      } else {
        reg[i & 2047] = acc;
        acc.x += 1;
      }
      i += 256;
    }
    if (acc.x == 0x12345678u) out[1] = acc.y;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int KIND, int KSTEPS>
static void run_conv(long long *out) {
  cudaFuncSetAttribute(kconv<KIND, KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  for (int N : {32, 64})
    for (int bg = 0; bg < 3; ++bg) {
      const int tiles = 128;
      for (int it = 0; it < 2; ++it) {
        kconv<KIND, KSTEPS><<<148, 320, 132 * 1024>>>(N, tiles, bg, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
      }
      printf("{\"pattern\": \"conv 9 taps x %d k-steps, %d MMAs per accumulator\", \"kind\": \"%s\", \"N\": %d, \"background\": \"%s\", \"cycles_per_mma\": %.1f, \"cycles_per_tap\": %.1f}\n",
             KSTEPS, 9 * KSTEPS, KIND ? "bf16" : "tf32", N, bg == 0 ? "none" : (bg == 1 ? "8 warps of 128-bit shared loads" : "8 warps of 128-bit shared stores"),
             (double)out[0] / (tiles * 9 * KSTEPS), (double)out[0] / (tiles * 9));
    }
}

int main() {
  long long *out;
  cudaMallocManaged(&out, 8);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 4096;
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {32, 64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > 512) continue;
        for (int variant = 0; variant < 2; ++variant) {   // 0: aligned operand, SBO 1024; 1: halo-style, SBO 1280, +384 B
          const uint32_t sbo = variant ? 1280 : 1024;
          const int a_off = variant ? 384 : 0;
          for (int it = 0; it < 2; ++it) {
            if (kind == 0) k<0><<<148, 128, 66 * 1024>>>(N, nacc, reps, sbo, a_off, out);
            else k<1><<<148, 128, 66 * 1024>>>(N, nacc, reps, sbo, a_off, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          printf("{\"kind\": \"%s\", \"N\": %d, \"accumulators\": %d, \"halo_desc\": %d, \"cycles_per_mma\": %.1f, \"floor\": %d}\n",
                 kind ? "bf16" : "tf32", N, nacc, variant, (double)out[0] / reps, N / 2);
        }
      }
  run_conv<0, 4>(out); run_conv<1, 4>(out);
  // fewer K-steps per tap: what one tap costs when only part of each 128-byte row is consumed
  run_conv<1, 2>(out); run_conv<1, 1>(out); run_conv<0, 2>(out); run_conv<0, 1>(out);
  return 0;
}
