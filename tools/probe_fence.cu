// Developer probe (not part of the product): what do concurrent PCIe copies do to the latency of
//   (a) a system-scope fence after a local / a peer store, (b) a device-scope fence, (c) an LL-packet ping-pong over NVLink?
// Motivation: bench.py's e2e leg at N >= 4 -- CG iterations run ~1.4x slower while H2D / D2H copies are in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/probe_fence tools/probe_fence.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ unsigned long long gns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// mode 0: store + fence.sys; 1: store + fence.gpu; 2: store only (volatile) ; 3: fence.sys without a store
__global__ void fence_kernel(volatile unsigned long long *target, int mode, unsigned long long run_ns, unsigned long long *out) {
    const unsigned long long t_begin = gns();
    unsigned long long n = 0, tot = 0, mx = 0;
    while (true) {
        const unsigned long long t0 = gns();
        if (t0 - t_begin > run_ns) break;
        if (mode != 3) *target = n;
        if (mode == 0 || mode == 3) __threadfence_system();
        else if (mode == 1) __threadfence();
        const unsigned long long t1 = gns();
        const unsigned long long d = t1 - t0;
        tot += d;
        if (d > mx) mx = d;
        ++n;
    }
    out[0] = n;
    out[1] = tot;
    out[2] = mx;
}

// LL ping-pong: `me` polls its own slot for epoch e, then stores e into the peer's slot.  rank 0 starts.
__global__ void pingpong_kernel(volatile unsigned long long *mine, volatile unsigned long long *theirs, int first, unsigned long long rounds,
                                unsigned long long *out) {
    unsigned long long mx = 0;
    const unsigned long long t_begin = gns();
    for (unsigned long long e = 1; e <= rounds; ++e) {
        const unsigned long long t0 = gns();
        if (first) *theirs = e;
        while (*mine < e) {
            if (gns() - t0 > 2000000000ull) { out[3] = 1; return; }
        }
        if (!first) *theirs = e;
        const unsigned long long d = gns() - t0;
        if (d > mx) mx = d;
    }
    out[0] = rounds;
    out[1] = gns() - t_begin;
    out[2] = mx;
}

int main(int argc, char **argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const size_t bytes = 256u << 20;
    const unsigned long long run_ns = 4000000ull;  // 4 ms per measurement; a 256 MiB copy takes ~5 ms
    void *h_in, *h_out;
    CK(cudaSetDevice(0));
    CK(cudaMallocHost(&h_in, bytes));
    CK(cudaMallocHost(&h_out, bytes));
    memset(h_in, 1, bytes);
    char *d_a, *d_b;
    CK(cudaMalloc(&d_a, bytes));
    CK(cudaMalloc(&d_b, bytes));
    unsigned long long *d_t, *d_out, h_res[4];
    CK(cudaMalloc(&d_t, 256));
    CK(cudaMalloc(&d_out, 256));
    CK(cudaMemset(d_t, 0, 256));
    cudaStream_t sk, s1, s2;
    CK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));

    unsigned long long *p_t = nullptr;  // a word in GPU 1's memory, mapped into GPU 0
    unsigned long long *p_out = nullptr;
    if (ndev > 1) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, 0, 1));
        if (can) {
            CK(cudaSetDevice(1));
            CK(cudaMalloc(&p_t, 256));
            CK(cudaMalloc(&p_out, 256));
            CK(cudaMemset(p_t, 0, 256));
            CK(cudaDeviceEnablePeerAccess(0, 0));
            CK(cudaSetDevice(0));
            CK(cudaDeviceEnablePeerAccess(1, 0));
        }
    }
    const char *copy_names[4] = {"idle", "H2D", "D2H", "H2D+D2H"};
    const char *mode_names[4] = {"store + fence.sys", "store + fence.gpu", "store only", "fence.sys alone"};
    for (int tgt = 0; tgt < (p_t ? 2 : 1); ++tgt) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int cp = 0; cp < 4; ++cp) {
                CK(cudaDeviceSynchronize());
                if (cp & 1) CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1));
                if (cp & 2) CK(cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2));
                fence_kernel<<<1, 1, 0, sk>>>(tgt ? p_t : d_t, mode, run_ns, d_out);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h_res, d_out, 32, cudaMemcpyDeviceToHost));
                printf("%-18s target %-5s copies %-8s: n %8llu  mean %8.2f us  max %8.2f us\n", mode_names[mode], tgt ? "peer" : "local", copy_names[cp],
                       h_res[0], h_res[0] ? h_res[1] / 1e3 / h_res[0] : 0.0, h_res[2] / 1e3);
            }
        }
    }
    if (p_t) {
        // LL ping-pong GPU0 <-> GPU1 under copies on GPU 0 (and on GPU 1: its own copies from a second pinned buffer)
        cudaStream_t sk1, s1b;
        void *h1;
        char *d1;
        CK(cudaSetDevice(1));
        CK(cudaStreamCreateWithFlags(&sk1, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s1b, cudaStreamNonBlocking));
        CK(cudaMallocHost(&h1, bytes));
        CK(cudaMalloc(&d1, bytes));
        for (int cp = 0; cp < 8; ++cp) {
            CK(cudaSetDevice(0));
            CK(cudaMemset(d_t, 0, 256));
            CK(cudaMemset(d_out, 0, 256));
            CK(cudaDeviceSynchronize());
            CK(cudaSetDevice(1));
            CK(cudaMemset(p_t, 0, 256));
            CK(cudaDeviceSynchronize());
            const unsigned long long rounds = 1500;
            CK(cudaSetDevice(0));
            if (cp & 1) CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1));
            if (cp & 2) CK(cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2));
            CK(cudaSetDevice(1));
            if (cp & 4) CK(cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, s1b));
            pingpong_kernel<<<1, 1, 0, sk1>>>(p_t, d_t, 0, rounds, p_out);
            CK(cudaSetDevice(0));
            pingpong_kernel<<<1, 1, 0, sk>>>(d_t, p_t, 1, rounds, d_out);
            CK(cudaDeviceSynchronize());
            CK(cudaSetDevice(1));
            CK(cudaDeviceSynchronize());
            CK(cudaSetDevice(0));
            CK(cudaMemcpy(h_res, d_out, 32, cudaMemcpyDeviceToHost));
            printf("LL ping-pong GPU0<->GPU1, copies on GPU0 %-8s%s: round trip mean %6.2f us  max %8.2f us%s\n", copy_names[cp & 3],
                   (cp & 4) ? " + D2H on GPU1" : "", h_res[1] / 1e3 / rounds, h_res[2] / 1e3, h_res[3] ? "  TIMEOUT" : "");
        }
    }
    return 0;
}
