// Microbenchmark: FP64 atomic throughput on B200 for the access patterns the J/K scatter can use.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash(unsigned x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
// mode 0: every lane a random address; mode 1: warp -> 32 consecutive doubles at random base;
// mode 2: groups of SEG consecutive doubles (lane/SEG group picks random base)
template<int MODE, int SEG>
__global__ void red_kernel(double* out, size_t n, int iters){
    unsigned tid = blockIdx.x*blockDim.x + threadIdx.x;
    unsigned lane = threadIdx.x & 31, warp = tid >> 5;
    for (int it = 0; it < iters; it++){
        size_t a;
        if (MODE == 0) a = hash(tid*977u + it*7919u) % n;
        else if (MODE == 1) a = (size_t)(hash(warp*977u + it*7919u) % (n/32))*32 + lane;
        else a = (size_t)(hash((tid/SEG)*977u + it*7919u) % (n/SEG))*SEG + (tid%SEG);
        atomicAdd(out + a, 1.0);
    }
}
__global__ void smem_kernel(double* out, int iters){
    __shared__ double s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 0;
    __syncthreads();
    unsigned tid = blockIdx.x*blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; it++) atomicAdd(&s[hash(tid*977u + it*7919u) & 4095], 1.0);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[1];
}
template<class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(){
    size_t sizes[2] = {size_t(1)<<21 /*16 MB*/, size_t(1)<<25 /*256 MB*/};
    int blocks = 148*8, threads = 256, iters = 256;
    double nops = (double)blocks*threads*iters;
    for (size_t n : sizes){
        double* d; cudaMalloc(&d, n*8); cudaMemset(d, 0, n*8);
        float t0 = timeit([&]{ red_kernel<0,1><<<blocks,threads>>>(d,n,iters); });
        float t1 = timeit([&]{ red_kernel<1,1><<<blocks,threads>>>(d,n,iters); });
        float t3 = timeit([&]{ red_kernel<2,3><<<blocks,threads>>>(d,n,iters); });
        float t4 = timeit([&]{ red_kernel<2,4><<<blocks,threads>>>(d,n,iters); });
        float t6 = timeit([&]{ red_kernel<2,6><<<blocks,threads>>>(d,n,iters); });
        float t8 = timeit([&]{ red_kernel<2,8><<<blocks,threads>>>(d,n,iters); });
        float t16 = timeit([&]{ red_kernel<2,16><<<blocks,threads>>>(d,n,iters); });
        printf("array %zu MB: scattered %.3e/s  coalesced32 %.3e/s  seg3 %.3e/s seg4 %.3e/s seg6 %.3e/s seg8 %.3e/s seg16 %.3e/s\n", n*8>>20,
               nops/t0*1e3, nops/t1*1e3, nops/t3*1e3, nops/t4*1e3, nops/t6*1e3, nops/t8*1e3, nops/t16*1e3);
        cudaFree(d);
    }
    double* o; cudaMalloc(&o, blocks*8);
    float ts = timeit([&]{ smem_kernel<<<blocks,threads>>>(o,iters); });
    printf("shared-memory atomicAdd(double) (CAS loop), 4096 slots: %.3e/s\n", nops/ts*1e3);
    return 0;
}
