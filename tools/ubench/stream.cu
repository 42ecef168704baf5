// Microbenchmark: achievable streaming-read rates at the tick's working-set sizes (50 MB .. 1 GB)
// with (a) plain 128-bit loads, (b) a TMA bulk-copy ring (the structure k_hist2 / k_place2 use).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, int hint) {
  if (hint) {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
  } else {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
  }
}
// (a) plain loads: each CTA owns a contiguous chunk; U independent 16B loads in flight per thread
template <int U>
__global__ void k_ldg(const int4* __restrict__ src, size_t n16, unsigned long long* sink, int contiguous) {
  size_t per = (n16 + gridDim.x - 1) / gridDim.x;
  size_t beg = contiguous ? per * blockIdx.x : (size_t)blockIdx.x * blockDim.x;
  size_t end = contiguous ? (beg + per < n16 ? beg + per : n16) : n16;
  size_t stride = contiguous ? blockDim.x : (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (size_t i = beg + (contiguous ? threadIdx.x : threadIdx.x); i < end; i += stride * U) {
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t j = i + u * stride; v[u] = j < end ? __ldcs(src + j) : make_int4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += (unsigned)v[u].x + (unsigned)v[u].w;
  }
  if (acc == 0x123456789ull) *sink = acc;
}
// (b) TMA ring: tile bytes TB, S stages, each CTA owns a contiguous chunk
__global__ void k_tma(const char* __restrict__ src, size_t nbytes, uint32_t TB, uint32_t S, int hint, unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)S * TB);
  size_t per = ((nbytes + gridDim.x - 1) / gridDim.x + TB - 1) / TB * TB;
  size_t beg = per * blockIdx.x, end = beg + per < nbytes ? beg + per : nbytes;
  if (beg >= nbytes) return;
  uint32_t nt = (uint32_t)((end - beg + TB - 1) / TB);
  if (threadIdx.x == 0) { for (uint32_t s = 0; s < S; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) for (uint32_t t = 0; t < S && t < nt; ++t) { mbar_expect(&full[t], TB); tma1d(sm + (size_t)t * TB, src + beg + (size_t)t * TB, TB, &full[t], hint); }
  unsigned long long acc = 0; uint32_t st = 0, par = 0;
  for (uint32_t t = 0; t < nt; ++t) {
    mbar_wait(&full[st], par);
    const int4* p = reinterpret_cast<const int4*>(sm + (size_t)st * TB);
    for (uint32_t i = threadIdx.x; i < TB / 16; i += blockDim.x) acc += (unsigned)p[i].x;
    __syncthreads();
    if (threadIdx.x == 0 && t + S < nt) { mbar_expect(&full[st], TB); tma1d(sm + (size_t)st * TB, src + beg + (size_t)(t + S) * TB, TB, &full[st], hint); }
    if (++st == S) { st = 0; par ^= 1; }
  }
  if (acc == 0x123456789ull) *sink = acc;
}
float timeit(void (*launch)(void*), void* ctx, void* flush, size_t fb) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
  for (int it = 0; it < 5; ++it) { cudaMemsetAsync(flush, it, fb); cudaEventRecord(a); launch(ctx); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms; }
  return best * 1000.f;
}
struct Ctx { const char* src; size_t n; int ctas, thr, U, contig; uint32_t TB, S; int hint; unsigned long long* sink; };
void l_ldg(void* c_) { Ctx* c = (Ctx*)c_; if (c->U == 4) k_ldg<4><<<c->ctas, c->thr>>>((const int4*)c->src, c->n / 16, c->sink, c->contig); else k_ldg<8><<<c->ctas, c->thr>>>((const int4*)c->src, c->n / 16, c->sink, c->contig); }
void l_tma(void* c_) { Ctx* c = (Ctx*)c_; size_t sm = (size_t)c->S * c->TB + 64; cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); k_tma<<<c->ctas, c->thr, sm>>>(c->src, c->n, c->TB, c->S, c->hint, c->sink); }
int main() {
  char* src; void* flush; unsigned long long* sink; size_t fb = 256u << 20;
  cudaMalloc(&src, 1u << 30); cudaMalloc(&flush, fb); cudaMalloc(&sink, 8); cudaMemset(src, 1, 1u << 30);
  for (size_t n : {(size_t)50 << 20, (size_t)100 << 20, (size_t)1 << 30}) {
    printf("--- read %zu MB\n", n >> 20);
    for (int contig : {0, 1}) for (int U : {4, 8}) for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
      Ctx c{src, n, ctas, 512, U, contig, 0, 0, 0, sink};
      float t = timeit(l_ldg, &c, flush, fb);
      printf("ldg128 %s U=%d ctas=%4d x512: %7.1f us  %6.2f TB/s\n", contig ? "contig-chunk" : "grid-stride ", U, ctas, t, n / t / 1e6);
    }
    for (uint32_t TB : {16384u, 32768u}) for (uint32_t S : {2u, 3u, 4u}) for (int ctas : {148, 296}) for (int hint : {0, 1}) {
      if ((size_t)S * TB * (ctas / 148) > 200 * 1024) continue;
      Ctx c{src, n, ctas, 512, 0, 1, TB, S, hint, sink};
      float t = timeit(l_tma, &c, flush, fb);
      printf("tma ring tile=%5u stages=%u ctas=%3d hint=%d: %7.1f us  %6.2f TB/s\n", TB, S, ctas, hint, t, n / t / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
