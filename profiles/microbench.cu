// Micro-benchmarks that decide the tile-kernel design on B200 (sm_100a): cost per warp-instruction of the candidate
// building blocks, measured with clock64() on one resident warp-set per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/microbench profiles/microbench.cu && profiles/microbench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE> __global__ void k_bench(unsigned long long *out_cycles, unsigned int *sink, int stride) {
    __shared__ unsigned int sm32[2048];
    __shared__ unsigned long long sm64[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm32[i] = i;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm64[i] = i;
    __syncthreads();
    unsigned int acc = threadIdx.x;
    float f = 1.0f + threadIdx.x * 1e-3f;
    unsigned long long a64 = threadIdx.x;
    const int warp_base = (threadIdx.x >> 5) * 64;
    long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) { // 32-bit smem atomic add, conflict-free (one bank per lane)
            atomicAdd(&sm32[warp_base + lane], acc);
        } else if (MODE == 1) { // 32-bit smem atomic add, pseudo-random addresses within 64 words
            atomicAdd(&sm32[warp_base + ((lane * stride + it) & 63)], acc);
        } else if (MODE == 2) { // 32-bit smem atomic with return (needed for carry)
            acc += atomicAdd(&sm32[warp_base + ((lane * stride + it) & 63)], acc);
        } else if (MODE == 3) { // 64-bit smem atomic add (CAS loop on this arch)
            atomicAdd(&sm64[(warp_base >> 1) + ((lane * stride + it) & 31)], a64);
        } else if (MODE == 4) { // shuffle
            acc = __shfl_sync(0xffffffffu, acc, (lane + 1) & 31);
        } else if (MODE == 5) { // LDS.32 conflict-free
            acc += sm32[warp_base + ((lane + acc) & 31)];
        } else if (MODE == 6) { // float -> int64 round-nearest
            a64 += static_cast<unsigned long long>(__float2ll_rn(f));
            f += 1.0f;
        } else if (MODE == 7) { // MUFU (ex2) chain-free
            f = __expf(f) * 1e-3f + 0.5f;
        } else if (MODE == 8) { // FFMA dependent chain (latency reference)
            f = fmaf(f, 1.0001f, 0.5f);
        } else if (MODE == 9) { // 64-bit shuffle
            a64 = __shfl_sync(0xffffffffu, a64, (lane + 1) & 31);
        } else if (MODE == 11) { // MATCH.ANY on a key with ~11 distinct values x multiplicity 3 (round-2 question: dedupe before ATOMS?)
            acc += __match_any_sync(0xffffffffu, (lane * stride + it + (acc & 1)) % 11);
        } else if (MODE == 12) { // REDUX.SUM over match groups (cooperative-groups labeled-partition pattern)
            const unsigned int m = 0x49249249u << (lane % 3); // three interleaved groups, precomputed mask
            acc += __reduce_add_sync(m, acc);
        } else if (MODE == 13) { // ATOMS.ADD.32, 3 lanes per address (the batch pattern of the tile kernel)
            atomicAdd(&sm32[warp_base + ((lane + it) % 11)], acc);
        } else if (MODE == 10) { // global RED.64 to warp-contiguous addresses (L2 atomics)
            atomicAdd(reinterpret_cast<unsigned long long *>(sink) + ((blockIdx.x * blockDim.x + threadIdx.x) & 8191), a64);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + static_cast<unsigned int>(f) + static_cast<unsigned int>(a64) + sm32[lane];
}

template <int MODE> void run(const char *name, int warps_per_block, int stride = 7) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long *d_cycles;
    unsigned int *d_sink;
    cudaMalloc(&d_cycles, sms * sizeof(unsigned long long));
    cudaMalloc(&d_sink, (size_t)sms * 1024 * sizeof(unsigned int) + 8192 * 8);
    cudaMemset(d_sink, 0, (size_t)sms * 1024 * sizeof(unsigned int) + 8192 * 8);
    k_bench<MODE><<<sms, warps_per_block * 32>>>(d_cycles, d_sink, stride);
    k_bench<MODE><<<sms, warps_per_block * 32>>>(d_cycles, d_sink, stride);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, d_cycles, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += h[i];
    avg /= sms;
    // cycles per warp-instruction as seen by the SM (all warps of the CTA issue ITERS each)
    printf("%-44s warps/SM=%2d  cycles/iter/warp=%8.2f  SM-cycles per warp-instr=%7.3f\n", name, warps_per_block,
           avg / ITERS, avg / ITERS / warps_per_block);
    cudaFree(d_cycles);
    cudaFree(d_sink);
}

int main() {
    for (int w : {8, 16}) {
        run<11>("MATCH.ANY (11 groups)", w);
        run<12>("REDUX.SUM (3 groups)", w);
        run<13>("ATOMS.ADD.32 3 lanes per address", w);
    }
    for (int w : {1, 8, 16}) {
        run<0>("ATOMS.ADD.32 conflict-free", w);
        run<1>("ATOMS.ADD.32 scattered(64 words)", w);
        run<2>("ATOMS.ADD.32 scattered, with return", w);
        run<3>("atomicAdd u64 smem (CAS loop)", w);
        run<4>("SHFL.32", w);
        run<9>("SHFL.64 (2x32)", w);
        run<5>("LDS.32", w);
        run<6>("F2I.S64.F32 (+IADD64)", w);
        run<7>("MUFU.EX2 (+FFMA)", w);
        run<8>("FFMA dependent chain", w);
        run<10>("RED.64 global, warp-contiguous", w);
        printf("\n");
    }
    return 0;
}
