// micro-benchmark: latency of cross-SM signalling through global memory (relaxed.gpu 16-byte states)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ ulonglong2 ld_state(const ulonglong2* p) {
    ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_state(ulonglong2* p, ulonglong2 v) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory"); }
__global__ void pingpong(ulonglong2* flags, int iters, long long* out) {
    // block 0 and block 1 bounce a counter
    if (threadIdx.x != 0) return;
    int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (me == 0) { st_state(&flags[0], make_ulonglong2(i, 0)); while (ld_state(&flags[8]).x != (unsigned long long)i) {} }
        else { while (ld_state(&flags[0]).x != (unsigned long long)i) {} st_state(&flags[8], make_ulonglong2(i, 0)); }
    }
    out[me] = clock64() - t0;
}
// chain: block b waits for flag[b-1] == i then sets flag[b] = i  (G blocks, one wave): measures per-hop latency
__global__ void chain(ulonglong2* flags, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    int b = blockIdx.x, G = gridDim.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (b > 0) { while (ld_state(&flags[b - 1]).x < (unsigned long long)i) {} }
        else if (i > 1) { while (ld_state(&flags[G - 1]).x < (unsigned long long)(i - 1)) {} }
        st_state(&flags[b], make_ulonglong2(i, 0));
    }
    out[b] = clock64() - t0;
}
// all-to-all style: every block publishes i, then waits until ALL lower blocks published i (32 lanes x 10 loads)
__global__ void lookback_like(ulonglong2* flags, int iters, long long* out) {
    int b = blockIdx.x, lane = threadIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (lane == 0) st_state(&flags[b], make_ulonglong2(i, 0));
        for (;;) {
            bool ok = true;
            for (int j = 0; j < 10; j++) { int idx = b - 1 - (lane * 10 + j); if (idx >= 0) ok = ok && (ld_state(&flags[idx]).x >= (unsigned long long)i); }
            if (__all_sync(0xffffffffu, ok)) break;
        }
    }
    if (lane == 0) out[b] = clock64() - t0;
}
int main() {
    ulonglong2* flags; long long* out; cudaMalloc(&flags, 4096 * 16); cudaMalloc(&out, 4096 * 8);
    long long h[512];
    int iters = 2000;
    cudaMemset(flags, 0, 4096 * 16);
    pingpong<<<2, 32>>>(flags, iters, out); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("pingpong: %.0f cycles per round trip (2 hops)\n", (double)h[0] / iters);
    for (int G : {8, 64, 148, 296}) {
        cudaMemset(flags, 0, 4096 * 16);
        chain<<<G, 32>>>(flags, 200, out); cudaMemcpy(h, out, G * 8, cudaMemcpyDeviceToHost);
        printf("chain G=%d: %.0f cycles per iteration, %.0f per hop\n", G, (double)h[G - 1] / 200, (double)h[G - 1] / 200 / G);
        cudaMemset(flags, 0, 4096 * 16);
        lookback_like<<<G, 32>>>(flags, iters, out); cudaMemcpy(h, out, G * 8, cudaMemcpyDeviceToHost);
        printf("lookback-like G=%d: %.0f cycles per iteration (last block)\n", G, (double)h[G - 1] / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
