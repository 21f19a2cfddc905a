// Data-parallel gradient exchange fused with the optimizer, over NVLink 5 / NVSwitch peer memory.
//
// The reference has no multi-GPU path; its only cross-replica operation is the FedAvg mean
// (Federated/utils.py:20-23).  Data parallelism here is "SUM the gradient slab over ranks, then Optimizer_Adam"
// (Adam.py:86-100).  A library all-reduce followed by an optimizer launch costs two dependent launches plus the
// collective's own latency for a 200 KB slab; this file does both in ONE kernel:
//
//   every rank owns a MAILBOX in its HBM, data[2 parities][world][slot] of 8-byte words {gradient float, step tag},
//   exported to the peers with cudaIpc.  A thread of the kernel owns four parameters:
//     1. it PUSHES its four local gradients, each packed with the step number into one 8-byte word (two words per
//        coalesced 16-byte store), into slot[rank] of every peer's mailbox (aligned 8-byte halves are single-copy atomic, so a word either carries this step's
//        tag AND this step's value or neither -- no fence, no separate flag: the low-latency protocol of collective
//        libraries),
//     2. it POLLS the same four words of every peer's slot in its own mailbox until their tags show this step,
//     3. it sums the `world` values in RANK ORDER (own value from the gradient slab; every rank computes the identical
//        fp32 sum, so the replicas stay bit-identical and the result does not depend on timing), writes the sum back to
//        the gradient slab and applies Adam to its parameters.
//   Threads only wait for pushes and push before they wait, so no rank can block another; the grid is capped at the SM
//   count so all blocks are resident.  Two parities make the buffers safe without a second handshake: a rank can only
//   be one step ahead of the slowest peer (it needs that peer's push to finish its own step), so a slot is never
//   overwritten while it may still be read.  The step number lives on the device and is advanced by the last block, so
//   the launch is CUDA-graph replayable.  A bounded spin (nm_dp_set_timeout_ms) raises a status word instead of hanging
//   the GPU if a peer never arrives.  Measured at 2 GPUs, 50 186 floats, back-to-back launches: see DESIGN.md section 5.
#include "nm_common.cuh"
#include "nm_adam.cuh"
#include <cstdlib>

namespace {

constexpr int kMaxWorld = 16;
constexpr int kChunk = 1024;                 // floats per chunk = 256 threads x float4
constexpr int kThreads = 256;

struct DpDev {                               // passed to the kernel by value
    uint2* mail[kMaxWorld];                  // mailbox of rank d ({value, tag} words), mapped into this process
    uint32_t* seq;                           // local: last completed step number
    uint32_t* ticket;                        // local: blocks finished in the current launch
    int* status;                             // local: 0 ok, 1 a wait timed out
    unsigned long long timeout_ns;
    size_t slot;                             // words per (parity, rank) slot, multiple of kChunk
    int rank, world;
    // two-phase mode (large slabs, world >= 4): reduce-scatter region [2 parities][world][gslot] followed by the all-gather
    // region [2 parities][aslot] in every mailbox
    int two_phase;
    size_t gslot, aslot;
};

struct Dp {
    DpDev dev{};
    void* base = nullptr;                    // this rank's mailbox allocation
    void* peer_base[kMaxWorld] = {};         // cudaIpcOpenMemHandle results (own entry = base)
    void* local = nullptr;                   // seq | ticket | status
    size_t data_bytes = 0, max_floats = 0;
    bool connected = false;
};

// two {value, tag} words in one 16-byte store: each aligned 8-byte half is single-copy atomic, and a warp's 32 stores
// cover 512 contiguous bytes (full sectors on the NVLink side: 8-byte stores at a 32-byte stride were packet-bound at 8 GPUs)
__device__ __forceinline__ void st_words(uint2* p, float v0, float v1, uint32_t tag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(v0)), "r"(tag),
                 "r"(__float_as_uint(v1)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_words(const uint2* p) {                          // two words; each half is self-validating
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Early push: chunks [ch_lo, ch_hi) of the gradient slab go to every peer's mailbox as soon as their producers have run
// (the backward pass launches this on a side stream behind each wgrad), so the exchange overlaps the rest of backward and
// the final kernel below finds them delivered.  Same words, same tags, same slots as the in-kernel push; stores only, no waits.
__global__ void __launch_bounds__(kThreads)
dp_push_kernel(DpDev c, const float* __restrict__ g, size_t n, int ch_lo, int ch_hi) {
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const uint32_t step = *reinterpret_cast<volatile uint32_t*>(c.seq) + 1u;
    const size_t par = (size_t)(step & 1u) * c.world;
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 7) == 0;
    for (int ch = ch_lo + blockIdx.x; ch < ch_hi; ch += gridDim.x) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t i = (size_t)ch * kChunk + h * (kChunk / 2) + (size_t)tid * 2;
            float2 L = make_float2(0.f, 0.f);
            if (vec && i + 2 <= n) L = *reinterpret_cast<const float2*>(g + i);
            else {
                if (i < n) L.x = g[i];
                if (i + 1 < n) L.y = g[i + 1];
            }
            for (int d = 1; d < c.world; ++d)
                st_words(c.mail[(c.rank + d) % c.world] + (par + c.rank) * c.slot + i, L.x, L.y, step);
        }
    }
}

// MODE 0: all-reduce only (gradient slab <- sum over ranks); MODE 1: + Adam on (p, m, v)
// Chunks >= pushed_from were already delivered by dp_push_kernel in this step: they are only collected here.
// A thread owns two pairs of parameters of its 1024-float chunk: floats {2t, 2t+1} and {512 + 2t, 512 + 2t + 1}.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
dp_allreduce_kernel(DpDev c, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    size_t n, nm_adam_state* ds, int n_apply, int advance, int pushed_from) {
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const uint32_t step = *reinterpret_cast<volatile uint32_t*>(c.seq) + 1u;
    const size_t par = (size_t)(step & 1u) * c.world;
    const int nchunks = (int)((n + kChunk - 1) / kChunk);
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 7) == 0;
    AdamHP hp{};
    if (MODE == 1) hp = adam_hp_from(ds);
    const uint2* mine = c.mail[c.rank];

    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        size_t idx[2];
        float2 L[2];                                                    // this rank's gradients (zero past the end)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t i = (size_t)ch * kChunk + h * (kChunk / 2) + (size_t)tid * 2;
            idx[h] = i;
            L[h] = make_float2(0.f, 0.f);
            if (vec && i + 2 <= n) L[h] = *reinterpret_cast<const float2*>(g + i);
            else {
                if (i < n) L[h].x = g[i];
                if (i + 1 < n) L[h].y = g[i + 1];
            }
        }
        // 1. push to every peer (destinations staggered by rank) -- unless dp_push_kernel already did
        for (int d = 1; d < c.world && ch < pushed_from; ++d) {
            uint2* q = c.mail[(c.rank + d) % c.world] + (par + c.rank) * c.slot;
            st_words(q + idx[0], L[0].x, L[0].y, step);
            st_words(q + idx[1], L[1].x, L[1].y, step);
        }
        // 2 + 3. collect the peers' values in rank order
        float2 G[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        for (int r = 0; r < c.world; ++r) {
            float2 t0 = L[0], t1 = L[1];
            if (r != c.rank) {
                const uint2* q = mine + (par + r) * c.slot;
                uint4 a = ld_words(q + idx[0]), b = ld_words(q + idx[1]);
                if (a.y != step || a.w != step || b.y != step || b.w != step) {
                    const unsigned long long t_start = global_ns();
                    unsigned polls = 0;
                    do {
                        if ((++polls & 255u) == 0 && global_ns() - t_start > c.timeout_ns) { atomicExch(c.status, 1); break; }
                        a = ld_words(q + idx[0]); b = ld_words(q + idx[1]);
                    } while (a.y != step || a.w != step || b.y != step || b.w != step);
                }
                t0 = make_float2(__uint_as_float(a.x), __uint_as_float(a.z));
                t1 = make_float2(__uint_as_float(b.x), __uint_as_float(b.z));
            }
            if (r == 0) { G[0] = t0; G[1] = t1; }
            else { G[0].x += t0.x; G[0].y += t0.y; G[1].x += t1.x; G[1].y += t1.y; }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t i = idx[h];
            if (vec && i + 2 <= n) {
                *reinterpret_cast<float2*>(g + i) = G[h];
                if (MODE == 1) {
                    float2 P = *reinterpret_cast<float2*>(p + i), M = *reinterpret_cast<float2*>(m + i),
                           V = *reinterpret_cast<float2*>(v + i);
                    adam_one(P.x, G[h].x, M.x, V.x, hp, n_apply); adam_one(P.y, G[h].y, M.y, V.y, hp, n_apply);
                    *reinterpret_cast<float2*>(p + i) = P; *reinterpret_cast<float2*>(m + i) = M;
                    *reinterpret_cast<float2*>(v + i) = V;
                }
            } else {
                const float gs[2] = {G[h].x, G[h].y};
                for (int u = 0; u < 2; ++u)
                    if (i + u < n) {
                        g[i + u] = gs[u];
                        if (MODE == 1) adam_one(p[i + u], gs[u], m[i + u], v[i + u], hp, n_apply);
                    }
            }
        }
    }

    // the last block to finish publishes the step number for the next launch (every block has read it by then)
    __syncthreads();
    if (tid == 0 && atomicAdd(c.ticket, 1u) == gridDim.x - 1) {
        *c.ticket = 0u;
        *reinterpret_cast<volatile uint32_t*>(c.seq) = step;
        if (MODE == 1 && advance) adam_state_advance(ds);      // every block has read the hyper-parameters by now
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-phase exchange for LARGE slabs (world >= 4): reduce-scatter + all-gather, O(n) traffic per rank instead of the
// O(world * n) of the all-to-all push above (VGG-style stack, 4.75 MB slab, 8 GPUs: 66 MB -> 16.6 MB pushed per rank and
// step).  Chunk c is OWNED by rank c % world.
//   push      a rank sends its gradients of chunk c only to the owner (same tagged 8-byte words, slot (c / world) of the
//             reduce-scatter region [parity][source rank]) -- dp_push2_kernel behind each wgrad, the rest by phase 0 below;
//   phase 1   the owner collects the world - 1 contributions, sums in RANK ORDER (bit-identical on every rank by
//             construction: only the owner forms the sum), and broadcasts the sum to every peer's all-gather region;
//   phase 2   every rank picks up the sums of the chunks it does not own.
// Every rank then holds the full summed gradient (as after an in-place all-reduce) and applies Adam to the whole slab, so
// parameters AND optimizer state stay replicated (checkpoints need no gather).  Waits only ever depend on pushes that are
// issued before any wait of the same kernel, so all-resident blocks cannot deadlock; the two-parity argument of the
// all-to-all protocol carries over (a rank needs every peer's step-s sums to finish step s).
__device__ __forceinline__ void load_local(const float* __restrict__ g, size_t n, bool vec, size_t i, float2& L) {
    L = make_float2(0.f, 0.f);
    if (vec && i + 2 <= n) L = *reinterpret_cast<const float2*>(g + i);
    else {
        if (i < n) L.x = g[i];
        if (i + 1 < n) L.y = g[i + 1];
    }
}

__device__ __forceinline__ uint4 poll_words(const uint2* q, uint32_t step, const DpDev& c) {
    uint4 a = ld_words(q);
    if (a.y != step || a.w != step) {
        const unsigned long long t_start = global_ns();
        unsigned polls = 0;
        do {
            if ((++polls & 255u) == 0 && global_ns() - t_start > c.timeout_ns) { atomicExch(c.status, 1); break; }
            a = ld_words(q);
        } while (a.y != step || a.w != step);
    }
    return a;
}

// reduce-scatter push of chunks [ch_lo, ch_hi): to the owner only
__global__ void __launch_bounds__(kThreads)
dp_push2_kernel(DpDev c, const float* __restrict__ g, size_t n, int ch_lo, int ch_hi) {
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const uint32_t step = *reinterpret_cast<volatile uint32_t*>(c.seq) + 1u;
    const size_t par = (size_t)(step & 1u);
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 7) == 0;
    for (int ch = ch_lo + blockIdx.x; ch < ch_hi; ch += gridDim.x) {
        const int owner = ch % c.world;
        if (owner == c.rank) continue;
        uint2* q = c.mail[owner] + (par * c.world + c.rank) * c.gslot + (size_t)(ch / c.world) * kChunk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t off = h * (kChunk / 2) + (size_t)tid * 2;
            float2 L;
            load_local(g, n, vec, (size_t)ch * kChunk + off, L);
            st_words(q + off, L.x, L.y, step);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
dp_allreduce2_kernel(DpDev c, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     size_t n, nm_adam_state* ds, int n_apply, int advance, int pushed_from) {
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const uint32_t step = *reinterpret_cast<volatile uint32_t*>(c.seq) + 1u;
    const size_t par = (size_t)(step & 1u);
    const int nchunks = (int)((n + kChunk - 1) / kChunk);
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 7) == 0;
    AdamHP hp{};
    if (MODE == 1) hp = adam_hp_from(ds);
    const uint2* my_rs = c.mail[c.rank] + par * c.world * c.gslot;
    const uint2* my_ag = c.mail[c.rank] + 2 * (size_t)c.world * c.gslot + par * c.aslot;

    auto finish = [&](size_t i, float2 G) {                   // summed gradient -> slab (+ Adam)
        if (vec && i + 2 <= n) {
            *reinterpret_cast<float2*>(g + i) = G;
            if (MODE == 1) {
                float2 P = *reinterpret_cast<float2*>(p + i), M = *reinterpret_cast<float2*>(m + i), V = *reinterpret_cast<float2*>(v + i);
                adam_one(P.x, G.x, M.x, V.x, hp, n_apply); adam_one(P.y, G.y, M.y, V.y, hp, n_apply);
                *reinterpret_cast<float2*>(p + i) = P; *reinterpret_cast<float2*>(m + i) = M; *reinterpret_cast<float2*>(v + i) = V;
            }
        } else {
            const float gs[2] = {G.x, G.y};
            for (int u = 0; u < 2; ++u)
                if (i + u < n) {
                    g[i + u] = gs[u];
                    if (MODE == 1) adam_one(p[i + u], gs[u], m[i + u], v[i + u], hp, n_apply);
                }
        }
    };

    // phase 0: reduce-scatter push of the chunks the backward pass has not delivered yet (the first layer's)
    const int lim = pushed_from < nchunks ? pushed_from : nchunks;
    for (int ch = blockIdx.x; ch < lim; ch += gridDim.x) {
        const int owner = ch % c.world;
        if (owner == c.rank) continue;
        uint2* q = c.mail[owner] + (par * c.world + c.rank) * c.gslot + (size_t)(ch / c.world) * kChunk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t off = h * (kChunk / 2) + (size_t)tid * 2;
            float2 L;
            load_local(g, n, vec, (size_t)ch * kChunk + off, L);
            st_words(q + off, L.x, L.y, step);
        }
    }
    // phase 1: owned chunks -- collect, sum in rank order, broadcast the sum
    for (int ch = c.rank + (int)blockIdx.x * c.world; ch < nchunks; ch += (int)gridDim.x * c.world) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t off = h * (kChunk / 2) + (size_t)tid * 2, i = (size_t)ch * kChunk + off;
            float2 L;
            load_local(g, n, vec, i, L);
            float2 G = make_float2(0.f, 0.f);
            for (int r = 0; r < c.world; ++r) {
                float2 t = L;
                if (r != c.rank) {
                    const uint4 a = poll_words(my_rs + (size_t)r * c.gslot + (size_t)(ch / c.world) * kChunk + off, step, c);
                    t = make_float2(__uint_as_float(a.x), __uint_as_float(a.z));
                }
                if (r == 0) G = t; else { G.x += t.x; G.y += t.y; }
            }
            for (int d = 1; d < c.world; ++d) {
                uint2* q = c.mail[(c.rank + d) % c.world] + 2 * (size_t)c.world * c.gslot + par * c.aslot + i;
                st_words(q, G.x, G.y, step);
            }
            finish(i, G);
        }
    }
    // phase 2: the other ranks' chunks -- pick up their sums
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        if (ch % c.world == c.rank) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t i = (size_t)ch * kChunk + h * (kChunk / 2) + (size_t)tid * 2;
            const uint4 a = poll_words(my_ag + i, step, c);
            finish(i, make_float2(__uint_as_float(a.x), __uint_as_float(a.z)));
        }
    }

    __syncthreads();
    if (tid == 0 && atomicAdd(c.ticket, 1u) == gridDim.x - 1) {
        *c.ticket = 0u;
        *reinterpret_cast<volatile uint32_t*>(c.seq) = step;
        if (MODE == 1 && advance) adam_state_advance(ds);
    }
}

int launch(Dp* d, int mode, float* p, float* g, float* m, float* v, size_t n, nm_adam_state* st, int n_apply, int advance, void* stream,
           int pushed_from = 0x7fffffff) {
    if (!d->connected) return nm::fail(NM_ERR_BAD_ARG, "%s%s", "nm_dp: not connected (call nm_dp_connect first)");
    if (n > d->max_floats) return nm::fail(NM_ERR_BAD_ARG, "%s%s", "nm_dp: slab larger than the mailbox slot");
    if (!n) return NM_OK;
    int dev = 0, sms = 0;
    NM_CUDA(cudaGetDevice(&dev));
    NM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int nchunks = (int)((n + kChunk - 1) / kChunk);
    const dim3 grid(nchunks < sms ? nchunks : sms);               // all blocks resident: a waiting block never starves a pusher
    if (d->dev.two_phase) {
        if (mode == 0)
            return nm::launch_pdl("dp_allreduce2", dp_allreduce2_kernel<0>, grid, dim3(kThreads), 0, nm::S(stream), d->dev,
                                  (float*)nullptr, g, (float*)nullptr, (float*)nullptr, n, (nm_adam_state*)nullptr, 0, 0, pushed_from);
        return nm::launch_pdl("dp_allreduce2_adam", dp_allreduce2_kernel<1>, grid, dim3(kThreads), 0, nm::S(stream), d->dev,
                              p, g, m, v, n, st, n_apply, advance, pushed_from);
    }
    if (mode == 0)
        return nm::launch_pdl("dp_allreduce", dp_allreduce_kernel<0>, grid, dim3(kThreads), 0, nm::S(stream), d->dev,
                              (float*)nullptr, g, (float*)nullptr, (float*)nullptr, n, (nm_adam_state*)nullptr, 0, 0, pushed_from);
    return nm::launch_pdl("dp_allreduce_adam", dp_allreduce_kernel<1>, grid, dim3(kThreads), 0, nm::S(stream), d->dev,
                          p, g, m, v, n, st, n_apply, advance, pushed_from);
}

}  // namespace

extern "C" {

int nm_dp_create(void** dp, int rank, int world, size_t max_floats) {
    NM_REQUIRE(dp, "dp is NULL");
    NM_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && max_floats > 0, "bad rank/world/size");
    Dp* d = new Dp();
    d->max_floats = max_floats;
    d->dev.rank = rank; d->dev.world = world;
    d->dev.slot = (max_floats + kChunk - 1) / kChunk * kChunk;
    d->dev.timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
    d->data_bytes = 2 * (size_t)world * d->dev.slot * sizeof(uint2);
    {
        // large slabs on >= 4 GPUs take the two-phase (reduce-scatter + all-gather) exchange: O(n) instead of O(world n) bytes
        // per rank; small slabs stay on the single-hop all-to-all (latency).  NM_DP_TWO_PHASE=0/1 overrides.
        const char* e = getenv("NM_DP_TWO_PHASE");
        d->dev.two_phase = e ? atoi(e) != 0 : (world >= 4 && max_floats >= (1u << 18));
        if (world < 2) d->dev.two_phase = 0;
        if (d->dev.two_phase) {
            const size_t nchunks = d->dev.slot / kChunk;
            d->dev.gslot = (nchunks + world - 1) / world * kChunk;
            d->dev.aslot = d->dev.slot;
            d->data_bytes = (2 * (size_t)world * d->dev.gslot + 2 * d->dev.aslot) * sizeof(uint2);
        }
    }
    cudaError_t e = cudaMalloc(&d->base, d->data_bytes);                     // cudaMalloc: cudaIpc cannot export pool memory
    if (e == cudaSuccess) e = cudaMemset(d->base, 0, d->data_bytes);         // tag 0 = "no step yet" (steps count from 1)
    if (e == cudaSuccess) e = cudaMalloc(&d->local, 256);
    if (e == cudaSuccess) e = cudaMemset(d->local, 0, 256);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d->base); cudaFree(d->local); delete d; return nm::cuda_fail(e, "nm_dp_create"); }
    d->dev.seq = static_cast<uint32_t*>(d->local);
    d->dev.ticket = d->dev.seq + 1;
    d->dev.status = reinterpret_cast<int*>(d->dev.seq + 2);
    d->peer_base[rank] = d->base;
    if (world == 1) {
        d->dev.mail[0] = static_cast<uint2*>(d->base);
        d->connected = true;
    }
    *dp = d;
    return NM_OK;
}

int nm_dp_export(void* dp, void* handle64) {
    NM_REQUIRE(dp && handle64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    NM_CUDA(cudaIpcGetMemHandle(&h, static_cast<Dp*>(dp)->base));
    memcpy(handle64, &h, sizeof(h));
    return NM_OK;
}

int nm_dp_connect(void* dp, const void* handles) {
    NM_REQUIRE(dp && handles, "null pointer");
    Dp* d = static_cast<Dp*>(dp);
    if (d->connected) return NM_OK;
    for (int r = 0; r < d->dev.world; ++r) {
        if (r != d->dev.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * sizeof(h), sizeof(h));
            NM_CUDA(cudaIpcOpenMemHandle(&d->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
        }
        d->dev.mail[r] = static_cast<uint2*>(d->peer_base[r]);
    }
    d->connected = true;
    return NM_OK;
}

int nm_dp_set_timeout_ms(void* dp, unsigned ms) {
    NM_REQUIRE(dp, "dp is NULL");
    static_cast<Dp*>(dp)->dev.timeout_ns = (unsigned long long)ms * 1000000ull;
    return NM_OK;
}

int nm_dp_allreduce_sum_f32(void* dp, float* buf, size_t n, void* stream) {
    NM_REQUIRE(dp && buf, "null pointer");
    return launch(static_cast<Dp*>(dp), 0, nullptr, buf, nullptr, nullptr, n, nullptr, 0, 0, stream);
}

int nm_dp_allreduce_adam_dev_f32(void* dp, float* param, float* grad, float* m, float* v, size_t n,
                                 nm_adam_state* dev_state, int n_apply, int advance, void* stream) {
    NM_REQUIRE(dp && param && grad && m && v && dev_state, "null pointer");
    NM_REQUIRE(n_apply >= 1 && n > 0, "n_apply must be >= 1 and n > 0");
    return launch(static_cast<Dp*>(dp), 1, param, grad, m, v, n, dev_state, n_apply, advance, stream);
}

int nm_dp_chunk_floats(void) { return kChunk; }

int nm_dp_push_f32(void* dp, const float* grad, size_t n, int chunk_lo, int chunk_hi, void* stream) {
    NM_REQUIRE(dp && grad, "null pointer");
    Dp* d = static_cast<Dp*>(dp);
    if (!d->connected) return nm::fail(NM_ERR_BAD_ARG, "%s%s", "nm_dp: not connected (call nm_dp_connect first)");
    if (n > d->max_floats) return nm::fail(NM_ERR_BAD_ARG, "%s%s", "nm_dp: slab larger than the mailbox slot");
    const int nchunks = (int)((n + kChunk - 1) / kChunk);
    NM_REQUIRE(chunk_lo >= 0 && chunk_lo <= chunk_hi && chunk_hi <= nchunks, "bad chunk range");
    if (chunk_lo == chunk_hi || d->dev.world == 1) return NM_OK;
    int dev = 0, sms = 0;
    NM_CUDA(cudaGetDevice(&dev));
    NM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // Grid size, measured at 8 GPUs (profiles/r02_dp_push_grid.txt): a SMALL slab (MNIST CNN: 49 chunks) is pushed right before
    // the collect kernel needs it, so it takes a block per chunk (4 blocks: 0.257 ms per step instead of 0.232); a LARGE
    // range (VGG: up to 577 chunks per layer) only has to land a backward pass later, and a 2-CTAs-per-SM grid took SM time
    // from the conv kernels it overlaps (1.848 ms per step against 1.821 with 16-64 blocks).  NM_DP_PUSH_BLOCKS overrides.
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("NM_DP_PUSH_BLOCKS"); forced = e ? atoi(e) : 0; }
    const int chunks = chunk_hi - chunk_lo;
    int blocks = forced > 0 ? forced : (chunks <= sms ? chunks : 64);
    if (blocks > chunks) blocks = chunks;
    if (d->dev.two_phase)
        return nm::launch_pdl("dp_push2", dp_push2_kernel, dim3(blocks), dim3(kThreads), 0, nm::S(stream), d->dev, grad, n, chunk_lo, chunk_hi);
    return nm::launch_pdl("dp_push", dp_push_kernel, dim3(blocks), dim3(kThreads), 0, nm::S(stream), d->dev, grad, n, chunk_lo, chunk_hi);
}

int nm_dp_collect_adam_dev_f32(void* dp, float* param, float* grad, float* m, float* v, size_t n,
                               nm_adam_state* dev_state, int n_apply, int advance, int pushed_from_chunk, void* stream) {
    NM_REQUIRE(dp && param && grad && m && v && dev_state, "null pointer");
    NM_REQUIRE(n_apply >= 1 && n > 0 && pushed_from_chunk >= 0, "n_apply must be >= 1, n > 0, pushed_from_chunk >= 0");
    return launch(static_cast<Dp*>(dp), 1, param, grad, m, v, n, dev_state, n_apply, advance, stream, pushed_from_chunk);
}

int nm_dp_status(void* dp, int* timed_out, unsigned* steps_done) {
    NM_REQUIRE(dp, "dp is NULL");
    Dp* d = static_cast<Dp*>(dp);
    uint32_t w[3];
    NM_CUDA(cudaMemcpy(w, d->local, sizeof(w), cudaMemcpyDeviceToHost));
    if (timed_out) *timed_out = (int)w[2];
    if (steps_done) *steps_done = w[0];
    return NM_OK;
}

int nm_dp_destroy(void* dp) {
    if (!dp) return NM_OK;
    Dp* d = static_cast<Dp*>(dp);
    cudaDeviceSynchronize();
    for (int r = 0; r < d->dev.world; ++r)
        if (r != d->dev.rank && d->peer_base[r]) cudaIpcCloseMemHandle(d->peer_base[r]);
    cudaFree(d->base);
    cudaFree(d->local);
    delete d;
    return NM_OK;
}

}  // extern "C"
