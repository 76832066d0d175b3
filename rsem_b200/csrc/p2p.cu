// Count-vector reduction over NVLink peer memory for the frozen-conprb rounds (N > 1 GPUs of one node).
//
// What it replaces: the reference's serial merge countvs[0][j] += countvs[i][j] (/root/reference/EM.cpp:385-389) -
// and, in this library, the ncclAllReduce of M + 1 doubles between K2 and K4.  A 1.6 MB allreduce is pure latency
// (~165 us measured at N = 8, against a K2 of ~0.25 ms per GPU when ONE C3 matrix is sharded over 8 GPUs), so the
// exchange is folded into the kernels of the round instead:
//   * every rank's K2 accumulates into one of two count buffers that all peers map (cudaIpc between processes,
//     peer access between the contexts of one process);
//   * the reduce kernel first tells every peer "my counts of round s are complete" (one system-scope release store
//     into the peer's flag array - K2 precedes it in stream order), waits until every peer has said the same, then each
//     rank sums the n buffers in rank order with independent coalesced loads over NVLink: every rank forms the same sum,
//     bit for bit, no broadcast, no second barrier.  While doing so it clears its own OTHER buffer for the next round
//     (peers read that one during round s - 1, which they finished before they signalled round s);
//   * K4 follows on the summed vector as in the single-GPU case.
// A rank whose peers never signal (a failed process) gives up after a few seconds and raises the context's error flag
// instead of spinning forever.  Model-updating rounds and the final pass keep ncclAllReduce (11 launches per run).
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace rsem_b200 {

namespace {

struct PeerInfo {   // exchanged with ncclAllGather
    cudaIpcMemHandle_t count_h, flag_h;
    long long pid;
    int device;
    int ok;
    void* count_ptr;
    void* flag_ptr;
};

struct ReduceArgs {
    const double* src[kMaxP2PRanks];          // every rank's count buffer of this round (own one included)
    unsigned long long* peer_flags[kMaxP2PRanks];
    const unsigned long long* my_flags;
    double* out;            // summed counts (K4's input)
    double* clear;          // own buffer of the next round
    int n_ranks, rank, M1;
    unsigned long long seq;
    int* err_flag;
    const int* done_flag;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) p2p_reduce_kernel(const ReduceArgs a) {
    __shared__ int s_fail;
    if (*a.done_flag) return;   // the loop already ended on every rank (same theta everywhere): nobody signals any more
    const int tid = threadIdx.x;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    if (blockIdx.x == 0 && tid < a.n_ranks && tid != a.rank) {
        __threadfence_system();
        st_release_sys(a.peer_flags[tid] + a.rank, a.seq);
    }
    if (tid < a.n_ranks && tid != a.rank) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(a.my_flags + tid) < a.seq) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > 5000000000ull) { s_fail = 1; break; }   // 5 s: a peer died
        }
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) *a.err_flag = 4;
        return;
    }
    const int stride = gridDim.x * blockDim.x * 2;
    for (int i = (blockIdx.x * blockDim.x + tid) * 2; i < a.M1; i += stride) {
        // two entries per thread, all ranks' loads issued before the first add: n independent NVLink reads in flight
        double v0[kMaxP2PRanks], v1[kMaxP2PRanks];
        const bool two = i + 1 < a.M1;
#pragma unroll
        for (int r = 0; r < kMaxP2PRanks; ++r) {
            v0[r] = r < a.n_ranks ? __ldcg(a.src[r] + i) : 0.0;
            v1[r] = (r < a.n_ranks && two) ? __ldcg(a.src[r] + i + 1) : 0.0;
        }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxP2PRanks; ++r) {   // rank order: the same sum on every rank
            s0 += v0[r];
            s1 += v1[r];
        }
        a.out[i] = s0;
        a.clear[i] = 0.0;
        if (two) { a.out[i + 1] = s1; a.clear[i + 1] = 0.0; }
    }
}

}  // namespace

void p2p_release(rsem_b200_ctx* c) {
    P2PState& p = c->p2p;
    for (int r = 0; r < p.n; ++r) {
        if (r == c->rank) continue;
        if (p.opened_count[r]) cudaIpcCloseMemHandle(p.opened_count[r]);
        if (p.opened_flags[r]) cudaIpcCloseMemHandle(p.opened_flags[r]);
    }
    if (p.count_buf) cudaFree(p.count_buf);
    if (p.flags) cudaFree(p.flags);
    p = P2PState{};
}

// Collective over the communicator: every rank calls it after its count vector size (M) is known.
// Leaves p.on == false (and the NCCL path in place) when any rank cannot map its peers.
int p2p_setup(rsem_b200_ctx* c) {
    P2PState& p = c->p2p;
    const bool keep_flags = p.flags != nullptr && p.n == c->n_ranks;
    unsigned long long* old_flags = keep_flags ? p.flags : nullptr;
    const unsigned long long old_seq = p.seq;
    if (!keep_flags) p2p_release(c);
    else {   // new matrix on the same communicator: only the count buffers change
        for (int r = 0; r < p.n; ++r)
            if (r != c->rank && p.opened_count[r]) { cudaIpcCloseMemHandle(p.opened_count[r]); p.opened_count[r] = nullptr; }
        if (p.count_buf) { cudaFree(p.count_buf); p.count_buf = nullptr; }
        p.on = false;
    }
    if (!c->comm || c->n_ranks < 2 || c->n_ranks > kMaxP2PRanks || getenv("RSEM_B200_NO_P2P")) return 0;
    const int n = c->n_ranks;
    const size_t M1 = (size_t)c->M + 1;
    PeerInfo mine;
    memset(&mine, 0, sizeof mine);
    mine.pid = (long long)getpid();
    mine.device = c->device;
    mine.ok = 1;
    if (cudaMalloc(&p.count_buf, 2 * M1 * sizeof(double)) != cudaSuccess) mine.ok = 0;
    if (mine.ok && !old_flags && cudaMalloc(&p.flags, kMaxP2PRanks * sizeof(unsigned long long)) != cudaSuccess) mine.ok = 0;
    if (mine.ok) {
        cudaMemsetAsync(p.count_buf, 0, 2 * M1 * sizeof(double), c->stream);
        if (!old_flags) cudaMemsetAsync(p.flags, 0, kMaxP2PRanks * sizeof(unsigned long long), c->stream);
        if (cudaIpcGetMemHandle(&mine.count_h, p.count_buf) != cudaSuccess) mine.ok = 0;
        if (cudaIpcGetMemHandle(&mine.flag_h, p.flags) != cudaSuccess) mine.ok = 0;
        mine.count_ptr = p.count_buf;
        mine.flag_ptr = p.flags;
    }
    cudaGetLastError();
    // exchange
    PeerInfo* d_all = nullptr;
    std::vector<PeerInfo> all(n);
    RB_CUDA(cudaMalloc(&d_all, (size_t)n * sizeof(PeerInfo)));
    RB_CUDA(cudaMemcpyAsync(d_all + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    if (int rc = nccl_allgather_bytes(c->comm, d_all + c->rank, d_all, sizeof(PeerInfo), c->stream)) { cudaFree(d_all); return rc; }
    RB_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)n * sizeof(PeerInfo), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_all);
    int ok = 1;
    for (int r = 0; r < n; ++r) ok &= all[r].ok;
    p.n = n;
    p.seq = keep_flags ? old_seq : 0;
    if (ok) {
        for (int r = 0; r < n && ok; ++r) {
            if (r == c->rank) { p.peer_count[r] = p.count_buf; p.peer_flags[r] = p.flags; continue; }
            if (all[r].pid == mine.pid) {   // another context of this process: plain peer access
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, all[r].device);
                if (!can && all[r].device != c->device) { ok = 0; break; }
                if (all[r].device != c->device) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                    cudaGetLastError();
                }
                p.peer_count[r] = static_cast<double*>(all[r].count_ptr);
                if (!keep_flags) p.peer_flags[r] = static_cast<unsigned long long*>(all[r].flag_ptr);
            } else {
                void* q = nullptr;
                if (cudaIpcOpenMemHandle(&q, all[r].count_h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
                p.opened_count[r] = q;
                p.peer_count[r] = static_cast<double*>(q);
                if (!keep_flags) {
                    if (cudaIpcOpenMemHandle(&q, all[r].flag_h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
                    p.opened_flags[r] = q;
                    p.peer_flags[r] = static_cast<unsigned long long*>(q);
                }
            }
        }
        cudaGetLastError();
    }
    // every rank must take the same path: agree on the minimum
    double* d_ok = nullptr;
    double h_ok = ok ? 1.0 : 0.0;
    RB_CUDA(cudaMalloc(&d_ok, sizeof(double)));
    RB_CUDA(cudaMemcpyAsync(d_ok, &h_ok, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (int rc = nccl_allreduce_sum_f64(c->comm, d_ok, 1, c->stream)) { cudaFree(d_ok); return rc; }
    RB_CUDA(cudaMemcpyAsync(&h_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_ok);
    p.on = h_ok > (double)n - 0.5;
    if (getenv("RSEM_B200_TIMING") && c->rank == 0)
        fprintf(stderr, "rsem_b200: count reduction over %s (%d ranks)\n", p.on ? "NVLink peer memory" : "ncclAllReduce", n);
    return 0;
}

// K2's accumulation target of the next frozen round
double* p2p_k2_target(rsem_b200_ctx* c) {
    P2PState& p = c->p2p;
    return p.count_buf + ((p.seq + 1) & 1ull) * ((size_t)c->M + 1);
}

// after K2 of a frozen round: barrier with the peers + sum of all ranks' buffers into c->count
int p2p_reduce(rsem_b200_ctx* c) {
    P2PState& p = c->p2p;
    const size_t M1 = (size_t)c->M + 1;
    p.seq++;
    ReduceArgs a;
    memset(&a, 0, sizeof a);
    for (int r = 0; r < p.n; ++r) {
        a.src[r] = p.peer_count[r] + (p.seq & 1ull) * M1;
        a.peer_flags[r] = p.peer_flags[r];
    }
    a.my_flags = p.flags;
    a.out = c->count;
    a.clear = p.count_buf + ((p.seq + 1) & 1ull) * M1;
    a.n_ranks = p.n;
    a.rank = c->rank;
    a.M1 = (int)M1;
    a.seq = p.seq;
    a.err_flag = c->err_flag;
    a.done_flag = c->done_flag;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)c->sm_count, (M1 / 2 + 255) / 256 + 1);
    p2p_reduce_kernel<<<grid, 256, 0, c->stream>>>(a);
    RB_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

}  // namespace rsem_b200
