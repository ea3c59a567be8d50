// Micro-benchmark: open-addressing hash aggregation of 64-bit keys into bucket-local tables (each bucket's
// table sized to stay in L2 while its keys stream through).  Decides whether a partition + hash-aggregate
// reduce could beat 6 radix passes + run-length reduce.   nvcc -arch=sm_100a -O3 -o hash_bench hash_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// keys of bucket b are i in [b*per, (b+1)*per); distinct keys per bucket = per * dup_frac
__global__ void insert_kernel(u64* tk, u32* tc, u64 n, u64 per, u64 slots_per_bucket, double distinct_frac) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 b = i / per;
        const u64 distinct = (u64)(per * distinct_frac) + 1;
        // zipf-ish: square of a uniform draw concentrates on small ids
        u64 r = mix(i) % distinct; r = (r * (mix(i * 31 + 7) % distinct)) / distinct;
        const u64 key = (b << 40) | (r + 1);
        u64 slot = mix(key) & (slots_per_bucket - 1);
        u64* bk = tk + b * slots_per_bucket; u32* bc = tc + b * slots_per_bucket;
        while (true) {
            const u64 old = atomicCAS(&bk[slot], 0ULL, key);
            if (old == 0ULL || old == key) { atomicAdd(&bc[slot], 1u); break; }
            slot = (slot + 1) & (slots_per_bucket - 1);
        }
    }
}
__global__ void compact_kernel(const u64* tk, const u32* tc, u64 slots, u32 min_count, u64* out_n) {
    u64 c = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += stride) c += (tk[i] != 0 && tc[i] >= min_count);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out_n, c);
}
int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 742000000ULL;
    const int n_buckets = argc > 2 ? atoi(argv[2]) : 256;
    const double load = argc > 3 ? atof(argv[3]) : 1.3;     // slots per key
    const u64 per = n / n_buckets;
    u64 slots = 1; while (slots < (u64)(per * load)) slots <<= 1;
    const u64 total = slots * n_buckets;
    u64 *tk, *out_n; u32* tc;
    cudaMalloc(&tk, total * 8); cudaMalloc(&tc, total * 4); cudaMalloc(&out_n, 8);
    cudaEvent_t e0, e1, e2, e3; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaMemsetAsync(tk, 0, total * 8); cudaMemsetAsync(tc, 0, total * 4); cudaMemsetAsync(out_n, 0, 8);
        cudaEventRecord(e1);
        insert_kernel<<<148 * 16, 256>>>(tk, tc, per * n_buckets, per, slots, 0.5);
        cudaEventRecord(e2);
        compact_kernel<<<148 * 16, 256>>>(tk, tc, total, 10, out_n);
        cudaEventRecord(e3);
        cudaEventSynchronize(e3);
        float a, b, c; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2); cudaEventElapsedTime(&c, e2, e3);
        u64 kept; cudaMemcpy(&kept, out_n, 8, cudaMemcpyDeviceToHost);
        printf("{\"n\": %llu, \"buckets\": %d, \"table_MB_per_bucket\": %.1f, \"memset_ms\": %.3f, \"insert_ms\": %.3f, \"compact_ms\": %.3f, \"kept\": %llu, \"err\": \"%s\"}\n",
               n, n_buckets, slots * 12 / 1e6, a, b, c, kept, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
