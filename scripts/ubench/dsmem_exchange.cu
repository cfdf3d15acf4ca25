// Microbenchmark for the K3 redesign question: how long does an all-to-all h exchange take inside a thread-block
// cluster when every CTA pushes its slice to every peer with st.async (data + mbarrier complete_tx in one
// instruction, no fences, no L2 round trip)?
//   cluster of CS CTAs; per iteration every CTA sends BYTES/CS bytes to each of the CS CTAs (itself included), so
//   every CTA receives BYTES per iteration; iteration i+1 starts when the local mbarrier of iteration i completed.
// Prints cycles per iteration (CTA 0 of cluster 0) and the number of co-resident clusters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_exchange dsmem_exchange.cu && ./dsmem_exchange
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async16(uint32_t dst, uint32_t mbar, uint4 v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(256) exch_kernel(int iters, int bytes, int tiles, long long *out, float *sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ uint64_t bar[4];                       // [tile][pingpong]
    cg::cluster_group cl = cg::this_cluster();
    const int rank = cl.block_rank();
    const int tid = threadIdx.x;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) for (int i = 0; i < 4; ++i) mbar_expect(&bar[i], bytes);
    cl.sync();
    const int per_dst = bytes / CS;                   // bytes this CTA sends to each destination
    const int atoms = per_dst / 16;
    long long t0 = clock64();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        for (int tl = 0; tl < tiles; ++tl) {
            const int pp = it & 1;
            unsigned char *buf = sm + (size_t)(tl * 2 + pp) * bytes;
            const uint32_t lb = smem_u32(buf) + rank * per_dst, lm = smem_u32(&bar[tl * 2 + pp]);
            const uint4 v = make_uint4(it, tid, rank, tl);
            for (int a = tid; a < atoms; a += blockDim.x)
#pragma unroll
                for (int d = 0; d < CS; ++d) st_async16(mapa(lb + a * 16, d), mapa(lm, d), v);
            // consume tile tl of this iteration
            mbar_wait(&bar[tl * 2 + pp], (it >> 1) & 1);
            acc += reinterpret_cast<const float *>(buf)[tid];
            __syncthreads();
            if (tid == 0) mbar_expect(&bar[tl * 2 + pp], bytes);      // re-arm for iteration it+2 (peers cannot be 2 ahead)
        }
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
    sink[blockIdx.x * blockDim.x + tid] = acc;
    cl.sync();
}

template <int CS>
static void run(int bytes, int tiles, int nclusters) {
    long long *out; float *sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 4 * 256 * 256);
    size_t smem = (size_t)tiles * 2 * bytes;
    cudaFuncSetAttribute(exch_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(exch_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * nclusters); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = -1;
    cudaOccupancyMaxActiveClusters(&maxc, exch_kernel<CS>, &cfg);
    int iters = 2000;
    cudaError_t e = cudaLaunchKernelEx(&cfg, exch_kernel<CS>, iters, bytes, tiles, out, sink);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("cluster %2d  clusters %d (max co-resident %d)  recv %6d B/tile x %d tiles: %lld cycles/iter  (%.1f B/cycle/SM in)  %s %s\n",
           CS, nclusters, maxc, bytes, tiles, h, (double)bytes * tiles / (double)h, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(out); cudaFree(sink);
}

int main() {
    run<10>(40960, 1, 1); run<10>(40960, 1, 8); run<10>(40960, 2, 8);
    run<10>(20480, 1, 8); run<10>(10240, 1, 8); run<10>(2560, 1, 8);
    run<8>(40960, 1, 8); run<8>(40960, 2, 8); run<8>(40960, 2, 16);
    run<15>(38400, 1, 4); run<15>(38400, 2, 8);
    run<16>(40960, 2, 8);
    return 0;
}
