// Probe for the multi-GPU gather design (tools only, not part of the library):
//   1. peer access + CUDA IPC memory handles between two PROCESSES (one per GPU, as under torchrun);
//   2. stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) on IPC-mapped peer memory as the
//      SM-free "shard arrived" / "slot consumed" signals;
//   3. bandwidth of SM stores into peer memory against a copy-engine cudaMemcpyAsync into the same mapping;
//   4. host memory placement: H2D bandwidth from page-locked memory bound (mbind) to each NUMA node.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o probe_p2p probe_p2p.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("[pid %d] %s -> %s (line %d)\n", getpid(), #x, cudaGetErrorString(e_), __LINE__); \
            fflush(stdout);                                                                     \
            _exit(3);                                                                           \
        }                                                                                       \
    } while (0)
#define CU(x)                                                                            \
    do {                                                                                 \
        CUresult r_ = (x);                                                               \
        if (r_ != CUDA_SUCCESS) {                                                        \
            const char* s_ = nullptr;                                                    \
            cuGetErrorString(r_, &s_);                                                   \
            printf("[pid %d] %s -> %s (line %d)\n", getpid(), #x, s_ ? s_ : "?", __LINE__); \
            fflush(stdout);                                                              \
            _exit(4);                                                                    \
        }                                                                                \
    } while (0)

__global__ void fill_kernel(uint4* dst, size_t n, unsigned v) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        dst[i] = make_uint4(v, v + 1, v + 2, unsigned(i));
}
__global__ void check_kernel(const uint4* src, size_t n, unsigned v, unsigned* bad) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        uint4 x = src[i];
        if (x.x != v || x.y != v + 1 || x.z != v + 2 || x.w != unsigned(i)) atomicAdd(bad, 1u);
    }
}

static const size_t BUF = size_t(64) << 20;  // payload bytes
static const size_t FLAG_OFF = BUF;          // flags live behind the payload in the same allocation

struct Msg {
    cudaIpcMemHandle_t h;
};

static void xwrite(int fd, const void* p, size_t n) {
    if (write(fd, p, n) != (ssize_t)n) _exit(5);
}
static void xread(int fd, void* p, size_t n) {
    size_t got = 0;
    while (got < n) {
        ssize_t k = read(fd, (char*)p + got, n - got);
        if (k <= 0) _exit(6);
        got += size_t(k);
    }
}

// rank 0 owns the receive buffer; rank 1 writes into it and signals
static int run_rank(int rank, int rfd, int wfd) {
    CK(cudaSetDevice(rank));
    CK(cudaFree(0));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    char* mine = nullptr;
    CK(cudaMalloc(&mine, BUF + 4096));
    CK(cudaMemset(mine, 0, BUF + 4096));
    CK(cudaDeviceSynchronize());
    Msg m{};
    CK(cudaIpcGetMemHandle(&m.h, mine));
    xwrite(wfd, &m, sizeof m);
    Msg peer{};
    xread(rfd, &peer, sizeof peer);
    char* theirs = nullptr;
    CK(cudaIpcOpenMemHandle((void**)&theirs, peer.h, cudaIpcMemLazyEnablePeerAccess));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, rank, 1 - rank));
    printf("[rank %d] IPC mapping ok, canAccessPeer=%d\n", rank, can);
    int memops = 0, flush = 0;
    CUdevice dev;
    CU(cuDeviceGet(&dev, rank));
    cuDeviceGetAttribute(&memops, CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_WAIT_VALUE_NOR, dev);
    cuDeviceGetAttribute(&flush, CU_DEVICE_ATTRIBUTE_CAN_FLUSH_REMOTE_WRITES, dev);
    printf("[rank %d] wait_value_nor=%d can_flush_remote_writes=%d\n", rank, memops, flush);
    fflush(stdout);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const size_t n16 = BUF / 16;
    if (rank == 1) {
        for (int round = 1; round <= 6; round++) {
            // flow control: wait until rank 0 has consumed the previous round (it writes round-1 into MY flag word)
            CU(cuStreamWaitValue32((CUstream)st, (CUdeviceptr)(mine + FLAG_OFF), unsigned(round - 1), CU_STREAM_WAIT_VALUE_GEQ));
            CK(cudaEventRecord(e0, st));
            const bool ce = round > 3;
            if (!ce) {
                fill_kernel<<<148 * 2, 512, 0, st>>>((uint4*)theirs, n16, unsigned(round * 1000));
            } else {
                fill_kernel<<<148 * 2, 512, 0, st>>>((uint4*)mine, n16, unsigned(round * 1000));
                CK(cudaMemcpyAsync(theirs, mine, BUF, cudaMemcpyDeviceToDevice, st));
            }
            CK(cudaEventRecord(e1, st));
            CU(cuStreamWriteValue32((CUstream)st, (CUdeviceptr)(theirs + FLAG_OFF), unsigned(round), CU_STREAM_WRITE_VALUE_DEFAULT));
            CK(cudaStreamSynchronize(st));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("[rank 1] round %d %s into peer memory: %.3f ms = %.1f GB/s\n", round, ce ? "local fill + copy-engine copy" : "SM stores",
                   ms, BUF / ms / 1e6);
            fflush(stdout);
        }
    } else {
        unsigned* bad = nullptr;
        CK(cudaMalloc(&bad, 4));
        for (int round = 1; round <= 6; round++) {
            CK(cudaMemsetAsync(bad, 0, 4, st));
            CU(cuStreamWaitValue32((CUstream)st, (CUdeviceptr)(mine + FLAG_OFF), unsigned(round), CU_STREAM_WAIT_VALUE_GEQ));
            check_kernel<<<148 * 2, 512, 0, st>>>((const uint4*)mine, n16, unsigned(round * 1000), bad);
            CU(cuStreamWriteValue32((CUstream)st, (CUdeviceptr)(theirs + FLAG_OFF), unsigned(round), CU_STREAM_WRITE_VALUE_DEFAULT));
            unsigned h = 99;
            CK(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            printf("[rank 0] round %d: payload seen after the flag, mismatching words = %u\n", round, h);
            fflush(stdout);
        }
    }
    CK(cudaDeviceSynchronize());
    CK(cudaIpcCloseMemHandle(theirs));
    return 0;
}

static long sys_mbind(void* addr, unsigned long len, int mode, const unsigned long* mask, unsigned long maxnode, unsigned flags) {
    return syscall(SYS_mbind, addr, len, mode, mask, maxnode, flags);
}

static void numa_probe(int n_dev) {
    const size_t bytes = size_t(1) << 30;
    for (int node = 0; node < 2; node++) {
        void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) {
            printf("mmap failed\n");
            return;
        }
        unsigned long mask = 1ul << node;
        long rc = sys_mbind(p, bytes, 2 /*MPOL_BIND*/, &mask, 64, 0);
        if (rc != 0) {
            printf("mbind(node %d) failed: %s\n", node, strerror(errno));
            munmap(p, bytes);
            continue;
        }
        memset(p, 1, bytes);
        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
            printf("cudaHostRegister failed on node %d\n", node);
            cudaGetLastError();
            munmap(p, bytes);
            continue;
        }
        for (int d = 0; d < n_dev; d++) {
            CK(cudaSetDevice(d));
            void* dptr;
            CK(cudaMalloc(&dptr, bytes));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            CK(cudaMemcpy(dptr, p, bytes, cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e0));
            for (int k = 0; k < 3; k++) CK(cudaMemcpyAsync(dptr, p, bytes, cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("H2D from NUMA node %d to GPU %d: %.1f GB/s\n", node, d, 3.0 * bytes / ms / 1e6);
            CK(cudaFree(dptr));
        }
        cudaHostUnregister(p);
        munmap(p, bytes);
    }
}

int main(int argc, char** argv) {
    int a2b[2], b2a[2];
    if (pipe(a2b) || pipe(b2a)) return 1;
    fflush(stdout);
    pid_t child = fork();  // before any CUDA call
    if (child == 0) {
        int rc = run_rank(1, a2b[0], b2a[1]);
        fflush(stdout);
        _exit(rc);
    }
    int rc = run_rank(0, b2a[0], a2b[1]);
    int status = 0;
    waitpid(child, &status, 0);
    printf("two-process IPC + stream memory operations: rank0 rc=%d rank1 status=%d\n", rc, status);
    if (argc > 1 && !strcmp(argv[1], "numa")) {
        int n = 0;
        CK(cudaGetDeviceCount(&n));
        numa_probe(n);
    }
    return 0;
}
