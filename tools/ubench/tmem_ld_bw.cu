// Micro-benchmark: tcgen05.ld throughput (TMEM -> registers) per SM as a function of the number of warps and the load shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench/tmem_ld_bw.cu -o gpurun_out/tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int SHAPE>   // 0: 32x32b.x32 (4 KB / warp instr), 1: 32x32b.x64 (8 KB), 2: 32x32b.x128 (16 KB), 3: 16x256b.x8 (4 KB; two per 32 lanes)
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, int nwarps, unsigned long long *out, float *sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = base + ((i * 32) & 255);
            if constexpr (SHAPE == 0) {
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(a));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[0] ^ r[31]);
            } else if constexpr (SHAPE == 1) {
                uint32_t r[64];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                               "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
                               "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                             : "r"(a & ~63u));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[0] ^ r[63]);
            } else {
                // two loads in flight before the wait (what the attention kernel does: chunks 1-3 issued under chunk 0)
                uint32_t r[32], q[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(a));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]),
                               "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                             : "r"((a + 32) & ~31u));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[0] ^ q[31]);
            }
        }
        t1 = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (lane == 0 && warp < nwarps) out[warp] = static_cast<unsigned long long>(t1 - t0);
    if (acc == 123.456f) sink[0] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr));
}

int main() {
    unsigned long long *d_out;
    float *d_sink;
    cudaMalloc(&d_out, 16 * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 4);
    const int iters = 4096;
    const char *names[3] = {"32x32b.x32 (1 in flight)", "32x32b.x64 (1 in flight)", "32x32b.x32 (2 in flight)"};
    const int bytes_per_iter[3] = {4096, 8192, 8192};
    for (int shape = 0; shape < 3; ++shape) {
        for (int nw : {1, 4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (shape == 0) tmem_ld_kernel<0><<<1, 512>>>(iters, nw, d_out, d_sink);
                if (shape == 1) tmem_ld_kernel<1><<<1, 512>>>(iters, nw, d_out, d_sink);
                if (shape == 2) tmem_ld_kernel<2><<<1, 512>>>(iters, nw, d_out, d_sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            unsigned long long h[16];
            cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
            unsigned long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            const double bytes = static_cast<double>(bytes_per_iter[shape]) * iters * nw;
            printf("%-28s %2d warps: %8llu cycles  -> %7.1f B/clk per SM, %6.1f cycles per warp instruction\n", names[shape], nw, mx, bytes / mx,
                   static_cast<double>(mx) / iters / (shape == 2 ? 2 : 1));
        }
    }
    return 0;
}
