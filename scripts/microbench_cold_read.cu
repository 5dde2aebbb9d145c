// Practical ceiling of a COLD streaming pass of the size k_step makes (three 4-byte-per-citizen streams + the count
// cells), timed the way bench.py times k_step: L2 flushed (256 MiB memset + read sweep), CUDA event, ONE launch, CUDA event.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o scripts/microbench_cold_read.bin scripts/microbench_cold_read.cu
//   ./scripts/microbench_cold_read.bin [citizens]
// Variants: plain (k_step's shape: 2 quads of 3 streams per thread and iteration, one wave of 148 x 4 blocks), the same
// with everything L2-prefetched up front, a deep-unroll single stream, and bulk asynchronous copies into shared memory.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_empty(uint32_t* sink) { if (sink == nullptr) return; }

__global__ void __launch_bounds__(256) k_sweep(const uint4* __restrict__ p, size_t n16, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldcg(p + i);
        acc ^= x.x ^ x.y ^ x.z ^ x.w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// k_step's access shape: 3 streams, 2 quads per thread and iteration
template <bool PF>
__global__ void __launch_bounds__(256, 4) k_three(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                                                  uint32_t n_quads, uint32_t* sink) {
    const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (PF && (threadIdx.x & 7u) == 0u) {
        for (uint32_t q = gtid + 2u * T; q < n_quads; q += T) { prefetch_l2(a + q); prefetch_l2(b + q); prefetch_l2(c + q); }
    }
    uint32_t acc = 0;
    for (uint32_t q0 = gtid; q0 < n_quads; q0 += 2u * T) {
        const uint32_t q1 = q0 + T;
        const bool h = q1 < n_quads;
        const uint4 z = make_uint4(0, 0, 0, 0);
        const uint4 wa = a[q0], wb = h ? a[q1] : z, ha = __ldg(b + q0), hb = h ? __ldg(b + q1) : z, ka = __ldg(c + q0), kb = h ? __ldg(c + q1) : z;
        acc ^= wa.x ^ wa.y ^ wa.z ^ wa.w ^ wb.x ^ wb.y ^ wb.z ^ wb.w ^ ha.x ^ ha.y ^ ha.z ^ ha.w ^ hb.x ^ hb.y ^ hb.z ^ hb.w ^
               ka.x ^ ka.y ^ ka.z ^ ka.w ^ kb.x ^ kb.y ^ kb.z ^ kb.w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;
}

// single stream, U independent 16-byte loads per thread in flight
template <int U>
__global__ void __launch_bounds__(256) k_unroll(const uint4* __restrict__ p, uint32_t n16, uint32_t* sink) {
    const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t i0 = gtid; i0 < n16; i0 += U * T) {
        uint4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t i = i0 + u * T; x[u] = i < n16 ? __ldg(p + i) : make_uint4(0, 0, 0, 0); }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= x[u].x ^ x[u].y ^ x[u].z ^ x[u].w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;
}

// bulk asynchronous copies (TMA, UBLKCP) into shared memory: one block per SM, STAGES tiles of TILE bytes in flight
constexpr int TILE = 16384, STAGES = 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k_bulk(const unsigned char* __restrict__ p, size_t bytes, uint32_t* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar[STAGES];
    const size_t per = ((bytes / gridDim.x + TILE - 1) / TILE) * TILE;
    const size_t lo = std::min(bytes, per * blockIdx.x), hi = std::min(bytes, lo + per);
    const uint32_t n_tiles = (uint32_t)((hi - lo + TILE - 1) / TILE);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t i) {
        const uint32_t s = i % STAGES;
        const size_t off = lo + (size_t)i * TILE;
        const uint32_t n = (uint32_t)std::min<size_t>(TILE, hi - off);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[s])), "r"(n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(sm + (size_t)s * TILE)), "l"(p + off), "r"(n), "r"(smem_u32(&bar[s])) : "memory");
    };
    if (threadIdx.x == 0) for (uint32_t i = 0; i < std::min<uint32_t>(STAGES, n_tiles); ++i) issue(i);
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n_tiles; ++i) {
        const uint32_t s = i % STAGES, parity = (i / STAGES) & 1u;
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
        } while (!done);
        acc ^= reinterpret_cast<const uint32_t*>(sm + (size_t)s * TILE)[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && i + STAGES < n_tiles) issue(i + STAGES);
    }
    if (acc == 0x9E3779B9u) *sink = acc;
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : 3452978u;
    const uint32_t n_quads = (n + 3) / 4;
    const size_t stream_bytes = (size_t)n_quads * 16, flush_bytes = (size_t)256 << 20;
    unsigned char *buf, *flush; uint32_t* sink;
    CK(cudaMalloc(&buf, 3 * stream_bytes + 1024)); CK(cudaMalloc(&flush, flush_bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, 3 * stream_bytes + 1024));
    const uint4* a = (const uint4*)buf; const uint4* b = a + n_quads; const uint4* c = b + n_quads;
    int sms = 148; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * STAGES));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double mb = 3.0 * stream_bytes / 1e6;
    printf("citizens %u: %.1f MB per pass, %d SMs\n", n, mb, sms);
    auto bench = [&](const char* name, auto launch, bool cold) {
        std::vector<float> t;
        for (int rep = 0; rep < 24; ++rep) {
            if (cold) {
                CK(cudaMemsetAsync(flush, rep & 0xFF, flush_bytes, st));
                k_sweep<<<sms * 8, 256, 0, st>>>((const uint4*)flush, flush_bytes / 16, sink);
            }
            CK(cudaEventRecord(e0, st));
            launch();
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep >= 4) t.push_back(ms * 1e3f);
        }
        std::sort(t.begin(), t.end());
        const float med = t[t.size() / 2];
        printf("%-34s %s  min %6.2f us  median %6.2f us  -> %7.0f GB/s (median)\n", name, cold ? "cold" : "warm", t[0], med, mb / med * 1e3);
    };
    for (int cold = 1; cold >= 0; --cold) {
        bench("empty kernel (event overhead)", [&] { k_empty<<<1, 32, 0, st>>>(sink); }, cold);
        bench("k_three 592x256 plain", [&] { k_three<false><<<sms * 4, 256, 0, st>>>(a, b, c, n_quads, sink); }, cold);
        bench("k_three 592x256 L2-prefetched", [&] { k_three<true><<<sms * 4, 256, 0, st>>>(a, b, c, n_quads, sink); }, cold);
        bench("k_three 1184x256 plain (2 waves)", [&] { k_three<false><<<sms * 8, 256, 0, st>>>(a, b, c, n_quads, sink); }, cold);
        bench("k_unroll<4> 1184x256", [&] { k_unroll<4><<<sms * 8, 256, 0, st>>>(a, 3 * n_quads, sink); }, cold);
        bench("k_unroll<8> 1184x256", [&] { k_unroll<8><<<sms * 8, 256, 0, st>>>(a, 3 * n_quads, sink); }, cold);
        bench("k_unroll<8> 592x256", [&] { k_unroll<8><<<sms * 4, 256, 0, st>>>(a, 3 * n_quads, sink); }, cold);
        bench("k_unroll<4> grid = n/1024", [&] { k_unroll<4><<<(3 * n_quads + 1023) / 1024, 256, 0, st>>>(a, 3 * n_quads, sink); }, cold);
        bench("k_bulk 148x128, 8 x 16 KB stages", [&] { k_bulk<<<sms, 128, TILE * STAGES, st>>>(buf, 3 * stream_bytes, sink); }, cold);
    }
    return 0;
}
