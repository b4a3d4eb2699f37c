// cuda_emu.h -- minimal CPU SIMT emulator used ONLY by the tests to execute the kernel source on the host.
//
// TEST INFRASTRUCTURE.  It exists because the build container has no GPU: the kernels in
// jadespectrogram_b200/csrc/jade_kernels.cuh are compiled a second time with -DJADE_EMU by g++ and run with one OS
// thread per CUDA thread (pthread barriers for __syncthreads/__syncwarp, a per-warp mailbox for shuffles), so that
// index arithmetic, shared-memory layouts and synchronisation can be debugged before GPU time is spent.
// Nothing in the product library includes this file and the product has no CPU code path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <vector>

struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace jade_emu {
struct BlockState {
    pthread_barrier_t block_bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<char> smem;
    std::vector<uint32_t> tmem = std::vector<uint32_t>(128 * 512, 0u); // tensor memory: [lane][column] words (jade_tmem.cuh)
    std::vector<float> mail; // one slot per thread: warp shuffles
    std::vector<unsigned long long> mail64;
    // named barriers (PTX bar.sync / bar.arrive id, count): arrivals counted per id, a generation per completed barrier
    std::mutex nb_mu;
    std::condition_variable nb_cv;
    unsigned nb_count[16] = {0};
    unsigned long long nb_gen[16] = {0};
};
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
// the block a thread belongs to (thread-local: the two CTAs of a cluster run at the same time, launch_cluster)
inline thread_local BlockState* g_block = nullptr;
inline void* dyn_smem() { return g_block->smem.data(); }
inline uint32_t* tmem() { return g_block->tmem.data(); }
// thread-block cluster of two CTAs: rank of this CTA, the peer's state, one barrier over all threads of the cluster
struct ClusterState {
    BlockState* cta[2] = {nullptr, nullptr};
    pthread_barrier_t bar;
};
inline thread_local ClusterState* t_cluster = nullptr;
inline thread_local unsigned t_cluster_rank = 0;
} // namespace jade_emu

#define threadIdx jade_emu::t_threadIdx
#define blockIdx jade_emu::t_blockIdx
#define blockDim jade_emu::g_blockDim
#define gridDim jade_emu::g_gridDim

inline void __syncthreads() { pthread_barrier_wait(&jade_emu::g_block->block_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&jade_emu::g_block->warp_bar[threadIdx.x >> 5]); }
// bar.arrive id, n: count this thread, do not wait.  bar.sync id, n: count this thread and wait until n threads have arrived.
inline void jade_emu_named_bar(int id, unsigned n, bool wait)
{
    jade_emu::BlockState& st = *jade_emu::g_block;
    std::unique_lock<std::mutex> lk(st.nb_mu);
    const unsigned long long gen = st.nb_gen[id];
    if (++st.nb_count[id] == n) {
        st.nb_count[id] = 0;
        ++st.nb_gen[id];
        st.nb_cv.notify_all();
        return;
    }
    if (wait) st.nb_cv.wait(lk, [&] { return st.nb_gen[id] != gen; });
}
inline float __shfl_xor_sync(unsigned, float v, int lane_mask)
{
    jade_emu::BlockState& st = *jade_emu::g_block;
    const unsigned tid = threadIdx.x;
    st.mail[tid] = v;
    __syncwarp();
    const float r = st.mail[(tid & ~31u) | ((tid ^ (unsigned)lane_mask) & 31u)];
    __syncwarp();
    return r;
}
inline unsigned long long __shfl_sync(unsigned, unsigned long long v, int src_lane)
{
    jade_emu::BlockState& st = *jade_emu::g_block;
    const unsigned tid = threadIdx.x;
    st.mail64[tid] = v;
    __syncwarp();
    const unsigned long long r = st.mail64[(tid & ~31u) | ((unsigned)src_lane & 31u)];
    __syncwarp();
    return r;
}
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline float sinpif(float x) { return (float)std::sin(M_PI * (double)x); }

namespace jade_emu {
// Launch kernel(args...) over grid x block threads; blocks run one after another.
template <typename K, typename... A>
void launch(K kernel, unsigned grid, unsigned block, size_t smem_bytes, A... args)
{
    g_blockDim = dim3(block);
    g_gridDim = dim3(grid);
    for (unsigned b = 0; b < grid; ++b) {
        BlockState st;
        st.smem.assign(smem_bytes + 64, 0);
        st.mail.assign(block, 0.f);
        st.mail64.assign(block, 0ull);
        pthread_barrier_init(&st.block_bar, nullptr, block);
        const unsigned nw = (block + 31) / 32;
        st.warp_bar.resize(nw);
        for (unsigned w = 0; w < nw; ++w) {
            unsigned cnt = (w + 1) * 32 <= block ? 32 : block - w * 32;
            pthread_barrier_init(&st.warp_bar[w], nullptr, cnt);
        }
        BlockState* stp = &st;
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([=]() {
                g_block = stp;
                t_threadIdx = dim3(t);
                t_blockIdx = dim3(b);
                kernel(args...);
            });
        for (auto& x : th) x.join();
        pthread_barrier_destroy(&st.block_bar);
        for (auto& wb : st.warp_bar) pthread_barrier_destroy(&wb);
    }
}

// The same for kernels launched as clusters of two CTAs (__cluster_dims__(2)): the two blocks of a cluster run concurrently
// and see each other's shared memory (cluster_peer_ptr) and a common barrier (cluster_sync).
template <typename K, typename... A>
void launch_cluster2(K kernel, unsigned grid, unsigned block, size_t smem_bytes, A... args)
{
    g_blockDim = dim3(block);
    g_gridDim = dim3(grid);
    for (unsigned c = 0; c + 1 < grid; c += 2) {
        BlockState st[2];
        ClusterState cl;
        pthread_barrier_init(&cl.bar, nullptr, 2 * block);
        const unsigned nw = (block + 31) / 32;
        for (int r = 0; r < 2; ++r) {
            st[r].smem.assign(smem_bytes + 64, 0);
            st[r].mail.assign(block, 0.f);
            st[r].mail64.assign(block, 0ull);
            pthread_barrier_init(&st[r].block_bar, nullptr, block);
            st[r].warp_bar.resize(nw);
            for (unsigned w = 0; w < nw; ++w) pthread_barrier_init(&st[r].warp_bar[w], nullptr, (w + 1) * 32 <= block ? 32 : block - w * 32);
            cl.cta[r] = &st[r];
        }
        ClusterState* clp = &cl;
        std::vector<std::thread> th;
        th.reserve(2 * block);
        for (unsigned r = 0; r < 2; ++r)
            for (unsigned t = 0; t < block; ++t)
                th.emplace_back([=]() {
                    g_block = clp->cta[r];
                    t_cluster = clp;
                    t_cluster_rank = r;
                    t_threadIdx = dim3(t);
                    t_blockIdx = dim3(c + r);
                    kernel(args...);
                });
        for (auto& x : th) x.join();
        for (int r = 0; r < 2; ++r) {
            pthread_barrier_destroy(&st[r].block_bar);
            for (auto& wb : st[r].warp_bar) pthread_barrier_destroy(&wb);
        }
        pthread_barrier_destroy(&cl.bar);
    }
}
inline unsigned cluster_rank() { return t_cluster_rank; }
inline void cluster_barrier() { pthread_barrier_wait(&t_cluster->bar); }
// address of the peer CTA's copy of a shared-memory object of this CTA
inline char* cluster_peer_ptr(const void* local, unsigned rank)
{
    const char* base = g_block->smem.data();
    return t_cluster->cta[rank]->smem.data() + (static_cast<const char*>(local) - base);
}
} // namespace jade_emu
