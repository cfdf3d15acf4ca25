// How long does MEMBAR.GPU (fence.acq_rel.gpu) take after a CTA's publish stores, as a function of the store shape?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membar_stores scripts/ubench/membar_stores.cu && /tmp/membar_stores
// 80 CTAs x 256 threads; per iteration every CTA writes 4 KB to rows 400 KB apart (the layer-output planes' geometry), then one
// thread fences and the clock around the fence is averaged.  mode 0: 8-byte pieces, two warps share a 32-byte sector (K3's
// publish today); mode 1: 16-byte pieces; mode 2: whole 32-byte sectors per lane pair; mode 3: whole 64-byte row segments by one
// quarter-warp; mode 4: no stores (fence alone).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k(unsigned char *buf, size_t row_stride, int iters, int mode, long long *out) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *base = buf + (size_t)blockIdx.x * 64;          // this CTA's 64-byte column window
    long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned char *plane = base + (size_t)(it & 1) * 8192;      // alternate two column windows
        const uint2 v2 = make_uint2(it, tid);
        const uint4 v4 = make_uint4(it, tid, it, tid);
        if (mode == 0) {
            // warp = (sp = warp & 3, half = warp >> 2): rows 16*half + (lane & 15), bytes 16*sp + 8*(lane >> 4); two planes
            const int row = 16 * (warp >> 2) + (lane & 15), off = 16 * (warp & 3) + 8 * (lane >> 4);
            *reinterpret_cast<uint2 *>(plane + row * row_stride + off) = v2;
            *reinterpret_cast<uint2 *>(plane + (32 + row) * row_stride + off) = v2;
        } else if (mode == 1) {
            const int row = 16 * (warp >> 2) + (lane & 15) + 32 * (lane >> 4), off = 16 * (warp & 3);
            *reinterpret_cast<uint4 *>(plane + row * row_stride + off) = v4;
        } else if (mode == 2) {
            // 8 warps x 32 lanes x 16 B = 4 KB: lane pair = one 32-byte sector; row = 8*warp + (lane >> 2), 64 rows
            const int row = 8 * warp + (lane >> 2), off = 16 * (lane & 3);
            *reinterpret_cast<uint4 *>(plane + row * row_stride + off) = v4;
        } else if (mode == 3) {
            const int row = 8 * warp + (lane >> 2), off = 16 * (lane & 3);
            *reinterpret_cast<uint4 *>(plane + row * row_stride + off) = v4;       // same as 2 (4 lanes = one 64-byte segment)
        }
        __syncthreads();
        if (tid == 0) {
            const long long t0 = clock64();
            __threadfence();
            const long long t1 = clock64();
            acc += t1 - t0;
        }
        // think time, as the rest of a step
        const long long w0 = clock64();
        while (clock64() - w0 < 6000) { }
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = acc / iters;
}

int main() {
    const size_t row_stride = 400640;      // T * Kpy * 2 bytes
    unsigned char *buf; long long *out;
    cudaMalloc(&buf, row_stride * 64 + (1 << 20));
    cudaMalloc(&out, 80 * sizeof(long long));
    for (int mode = 0; mode <= 4; ++mode) {
        if (mode == 3) continue;
        k<<<80, 256>>>(buf, row_stride, 300, mode, out);
        cudaDeviceSynchronize();
        long long h[80];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long s = 0, mx = 0;
        for (int i = 0; i < 80; ++i) { s += h[i]; if (h[i] > mx) mx = h[i]; }
        printf("mode %d: fence cycles mean %lld max %lld  (%s)\n", mode, s / 80, mx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
