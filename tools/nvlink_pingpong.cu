// One-way latency of the primitives the in-kernel rank exchange is built from, measured between two B200s over NVLink:
// a single warp on GPU 0 and a single warp on GPU 1 bounce a sequence number through peer-mapped memory.
//   st   : st.relaxed.sys.u64 into the peer's word, the peer polls its own memory with ld.relaxed.sys
//   red  : red.relaxed.sys.add.u64 into the peer's word (the one-hop / hierarchical sums), same poll
//   st8  : eight stores to eight different peers' words are modelled by 8 stores to 8 lines of the one peer
// Also the local legs on one GPU: red.gpu + poll by another SM, atom.gpu (returning) round trip.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/nvlink_pingpong tools/nvlink_pingpong.cu
//   (needs 2 GPUs: gpurun --gpus 2)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("%s: %s\n", #x, cudaGetErrorString(e_));                            \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// mode 0: st, 1: red (adds 1 per round: the word counts rounds), 2: st to 8 lines (only line 0 is polled)
__global__ void bounce(int me, int mode, int rounds, unsigned long long *mine, unsigned long long *peer, long long *cycles) {
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    for (int i = 1; i <= rounds; ++i) {
        if (me == 0) {  // initiator: send i, wait for the echo
            if (mode == 1)
                red_sys(peer, 1ull);
            else {
                if (mode == 2)
                    for (int q = 7; q >= 1; --q) st_sys(peer + 16 * q, (unsigned long long)i);
                st_sys(peer, (unsigned long long)i);
            }
            while (ld_sys(mine) < (unsigned long long)i) {
            }
        } else {  // responder
            while (ld_sys(mine) < (unsigned long long)i) {
            }
            if (mode == 1)
                red_sys(peer, 1ull);
            else {
                if (mode == 2)
                    for (int q = 7; q >= 1; --q) st_sys(peer + 16 * q, (unsigned long long)i);
                st_sys(peer, (unsigned long long)i);
            }
        }
    }
    if (me == 0) *cycles = clock64() - t0;
}

// local legs on one GPU: CTA 0 and CTA 1 (different SMs) bounce through L2
__global__ void local_bounce(int mode, int rounds, unsigned long long *a, unsigned long long *b, long long *cycles) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    unsigned long long *mine = me == 0 ? a : b, *peer = me == 0 ? b : a;
    const long long t0 = clock64();
    for (int i = 1; i <= rounds; ++i) {
        if (me == 0) {
            if (mode == 1)
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(peer), "l"(1ull) : "memory");
            else
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(peer), "l"((unsigned long long)i) : "memory");
            unsigned long long v;
            do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            } while (v < (unsigned long long)i);
        } else {
            unsigned long long v;
            do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            } while (v < (unsigned long long)i);
            if (mode == 1)
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(peer), "l"(1ull) : "memory");
            else
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(peer), "l"((unsigned long long)i) : "memory");
        }
    }
    if (me == 0) *cycles = clock64() - t0;
}

__global__ void atom_rt(int rounds, unsigned long long *a, long long *cycles) {
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    unsigned long long acc = 0;
    for (int i = 0; i < rounds; ++i) {
        unsigned long long old;
        asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(a), "l"(1ull + (acc & 0)) : "memory");
        acc += old;  // dependent chain
    }
    *cycles = clock64() - t0 + (long long)(acc & 0);
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    const int rounds = 2000;
    {  // local legs
        CK(cudaSetDevice(0));
        unsigned long long *a, *b;
        long long *cyc, h = 0;
        CK(cudaMalloc(&a, 4096));
        CK(cudaMalloc(&b, 4096));
        CK(cudaMalloc(&cyc, 8));
        for (int mode = 0; mode < 2; ++mode) {
            CK(cudaMemset(a, 0, 4096));
            CK(cudaMemset(b, 0, 4096));
            void *args[] = {(void *)&mode, (void *)&rounds, (void *)&a, (void *)&b, (void *)&cyc};
            CK(cudaLaunchCooperativeKernel((const void *)local_bounce, dim3(2), dim3(32), args, 0, 0));
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("one GPU, SM <-> SM through L2, %-4s: %6.0f cycles round trip (%4.0f one way)\n", mode ? "red" : "st",
                   (double)h / rounds, (double)h / rounds / 2);
        }
        CK(cudaMemset(a, 0, 4096));
        atom_rt<<<1, 32>>>(rounds, a, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("one GPU, atom.add.u64 returning the old value: %6.0f cycles per dependent atomic\n", (double)h / rounds);
    }
    if (n < 2) {
        printf("one GPU only: NVLink legs skipped\n");
        return 0;
    }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (!can) {
        printf("no peer access between GPU 0 and 1\n");
        return 0;
    }
    unsigned long long *buf[2];
    long long *cyc;
    cudaStream_t st[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&buf[d], 4096));
        CK(cudaStreamCreate(&st[d]));
    }
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&cyc, 8));
    const char *names[] = {"st.relaxed.sys", "red.relaxed.sys.add", "8 x st (8 lines)"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaMemset(buf[d], 0, 4096));
            CK(cudaDeviceSynchronize());
        }
        for (int d = 1; d >= 0; --d) {  // responder first
            CK(cudaSetDevice(d));
            bounce<<<1, 32, 0, st[d]>>>(d, mode, rounds, buf[d], buf[1 - d], cyc);
        }
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaDeviceSynchronize());
        }
        long long h = 0;
        CK(cudaSetDevice(0));
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("GPU 0 <-> GPU 1 over NVLink, %-20s: %6.0f cycles round trip = %.2f us one way (at 1.965 GHz)\n", names[mode],
               (double)h / rounds, (double)h / rounds / 2 / 1965.0);
    }
    return 0;
}
