// pv_kernels.cu -- sm_100a kernels + the C-ABI of include/panda_validity.h (state / edge validity, FK,
// device-generated sweeps, FP32 peak probe).  The batched RRT-Connect lives in pv_rrtc.cu.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/panda_validity.h"
#include "pv_device.cuh"
#include "pv_handle.h"

#ifndef PV_THREADS
#define PV_THREADS 128
#endif
#ifndef PV_MIN_BLOCKS
#define PV_MIN_BLOCKS 3
#endif

static char g_create_error[512] = "";

#define PV_CUDA(h, expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf((h)->err, sizeof((h)->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                             \
            return PV_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

// =========================================================================================================
// kernels
// =========================================================================================================

// (pv_emit_word, the epilogue shared by the verdict-producing kernels, lives in pv_handle.h)

// K2: one thread per configuration, warp-ballot bit packing.  Persistent blocks; the warps of a block walk the
// verdict words in lockstep (block barriers inside pv_check_config) so they share instruction fetches.
#ifndef PV_SB_THREADS
#define PV_SB_THREADS 512  // 16 warps in lockstep, 127 registers, no spills (swept 128..1024: profiles/r1_notes.md)
#endif
// extra block barriers INSIDE the check (levels 1..3 of pv_check_config's SYNC); 0 = only the one per iteration below,
// which measured best (profiles/r1_notes.md)
#ifndef PV_SB_SYNC
#define PV_SB_SYNC 0
#endif
// Lockstep of the sorted kernel's main loop: 1 = one block barrier per iteration (what made the 60 KB loop of round 1
// share its instruction fetches), n > 1 = named barriers inside n groups of warps, 0 = none.  Since the scene section left
// the loop (PV_COLD_SCENE) and the general sphere-vs-box tests left the kernel (YAW), the loop is 27 KB and fits the
// instruction cache: free-running warps no longer miss, and the barrier -- the top stall, 1.39 warp-cycles per issue -- only
// made every warp wait for the slowest.  Same box, resident 2 Mi batch: goal 1 22.6 -> 23.8, tower 16.4 -> 16.9, pentagon
// 21.3 -> 22.1 G checks/s (groups of 4 warps: 23.1 / 16.4 / 21.5).
#ifndef PV_SB_BAR_GROUPS
#define PV_SB_BAR_GROUPS 0
#endif
#ifndef PV_OWED_SPLIT
#define PV_OWED_SPLIT 1  // owed scene sections dealt out box-wise to idle lanes (sorted state kernel)
#endif
#ifndef PV_SB_MINB
#define PV_SB_MINB 1
#endif
template <bool AOS, bool CULL, bool CARRY>
__global__ void __launch_bounds__(PV_SB_THREADS, PV_SB_MINB)
    pv_state_bits_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ qA,
                         const float4* __restrict__ qB, const float* __restrict__ q9,
                         const float* __restrict__ q_aos, int64_t n, uint32_t* __restrict__ bits,
                         const __grid_constant__ PvGather G) {
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t n_words = (n + 31) >> 5;
    // block-uniform trip count: every warp of the block runs every iteration (barriers inside)
    for (int64_t wb = (int64_t)blockIdx.x * warps_per_block; wb < n_words; wb += (int64_t)gridDim.x * warps_per_block) {
        const int64_t w = wb + warp_in_block;
        const int64_t i = (w << 5) + lane;
        const bool in = i < n;
        const int64_t ii = in ? i : n - 1;
        float q[9];
        if (AOS) pv_load_aos(q_aos, ii, q);
        else pv_load_soa(qA, qB, q9, ii, q);
        __syncthreads();  // lockstep: the 16 warps of the block share instruction fetches
        PvAcc<PV_MODE_BITS> acc;
        pv_check_config<PV_MODE_BITS, CULL, PV_EXIT_NONE, PV_SB_SYNC, true, CARRY>(q, S, acc);
        const unsigned word = __ballot_sync(0xffffffffu, in && !acc.hit);
        if (w < n_words) pv_emit_word(bits, G, w, word, lane);
    }
}

// ---- sorted variant ------------------------------------------------------------------------------------------
// The per-lane culls only pay when a whole warp can skip a block of tests, and with unordered inputs some lane of the
// 32 nearly always needs it (self-collision blocks ran with ~8 of 32 lanes live: profiles/r1_notes.md).  Which blocks
// a configuration needs is decided mostly by the elbow angle q[3] (it alone fixes the shoulder-wrist distance), the wrist
// flex q[5] and how far the wrist is from the scene boxes (pv_sort_key), so a block first counting-sorts ITS share of
// the batch (the same 512-configuration chunks the unsorted kernel gives it, up to PV_ST configurations at a time) by
// that key in shared memory -- keys and 16-bit indices only; the configurations stay in global memory / L2 and are
// gathered -- and then walks that share in sorted order: the 32 lanes of a warp, and the 16 warps that meet at the
// lockstep barrier, hold near-equal keys and skip or take the same blocks.
// Verdict bits return to their ORIGINAL positions through a shared-memory bit array, so the output (and the fused
// gather) is unchanged.
#ifndef PV_ST
#define PV_ST 16384  // super-tile: at most this many configurations are sorted together (32 chunks)
#endif
#define PV_SORT_BUCKETS 256
// key loads a thread keeps in flight in the key pass (where the batch is first read from HBM)
#ifndef PV_KEY_LOADS
#define PV_KEY_LOADS 4  // swept 1 / 2 / 4 / 8 / 14 under bench.py: 14.83 / 15.50 / 15.57 / 15.38 / 15.2 G checks/s
#endif
static_assert(PV_ST % PV_SB_THREADS == 0 && PV_ST <= 65536, "super-tile = whole chunks, indices fit 16 bits");

__device__ __forceinline__ void pv_emit_word_thread(uint32_t* __restrict__ bits, const PvGather& G, int64_t w, unsigned word) {
    if (bits) bits[w] = word;
    if (w >= G.word_cap) return;
    if (G.mc) {
        asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(G.mc + G.word_off + w), "r"(word) : "memory");
    } else if (G.peers) {
        for (int p = 0; p < G.n_peers; ++p) G.peers[p][G.word_off + w] = word;
    }
}

// Sort key = (distance class of the wrist point from the scene's bounds, elbow-angle bin).  The elbow angle decides
// most self-collision culls; how far the wrist (origin of link5/link6, a function of q[0..3] only) is from the bounds of
// the scene boxes decides whether the scene-level cull of pv_check_config lets the whole warp skip the box loop.  The key
// only orders the work (verdicts return to their original positions), so it is computed with the fast intrinsics.
#ifndef PV_SORT_E0
#define PV_SORT_E0 0.20f
#define PV_SORT_E1 0.35f
#define PV_SORT_E2 0.50f
#endif
// the wrist-flex angle q[5] decides the forearm-vs-finger block, the most frequent self-collision block left
#ifndef PV_SORT_Q5_BINS
#define PV_SORT_Q5_BINS 4
#endif
#define PV_SORT_Q3_BINS (PV_SORT_BUCKETS / 4 / PV_SORT_Q5_BINS)
__device__ __forceinline__ int pv_sort_key(float q0, float q1, float q2, float q3, float q5, const PvScene& S) {
    const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
    const float scale = (float)PV_SORT_Q3_BINS / (hi[3] - lo[3]);
    int bin = (int)fminf(fmaxf((q3 - lo[3]) * scale, 0.f), (float)(PV_SORT_Q3_BINS - 1));  // NaN -> bin 0
    if (PV_SORT_Q5_BINS > 1) {
        const float scale5 = (float)PV_SORT_Q5_BINS / (hi[5] - lo[5]);
        bin = bin * PV_SORT_Q5_BINS + (int)fminf(fmaxf((q5 - lo[5]) * scale5, 0.f), (float)(PV_SORT_Q5_BINS - 1));
    }
    float s, c;
    __sincosf(q0, &s, &c);
    float3 p = make_float3(S.base[0], S.base[1], S.base[2] + 0.333f);
    float3 X = make_float3(c, s, 0.f), Y = make_float3(-s, c, 0.f), Z, Xp, Yp, Zp;
    __sincosf(q1, &s, &c);  // link2: Rx(-90)
    {
        const float3 X1 = X;
        X = make_float3(c * X1.x, c * X1.y, -s);
        Z = Y;
        Y = make_float3(-s * X1.x, -s * X1.y, -c);
    }
    __sincosf(q2, &s, &c);  // link3: pos (0,-0.316,0), Rx(+90)
    p = v_fma(Y, -0.316f, p);
    Xp = X; Yp = Z; Zp = v_neg(Y);
    v_rotz(Xp, Yp, c, s, X, Y); Z = Zp;
    __sincosf(q3, &s, &c);  // link4: pos (0.0825,0,0), Rx(+90)
    p = v_fma(X, 0.0825f, p);
    Xp = X; Yp = Z;
    v_rotz(Xp, Yp, c, s, X, Y);
    p = v_fma(Y, 0.384f, v_fma(X, -0.0825f, p));  // link5 origin: pos (-0.0825,0.384,0) in link4
    const float dx = fmaxf(fmaxf(S.aabb_lo[0] - p.x, p.x - S.aabb_hi[0]), 0.f);
    const float dy = fmaxf(fmaxf(S.aabb_lo[1] - p.y, p.y - S.aabb_hi[1]), 0.f);
    const float dz = fmaxf(fmaxf(S.aabb_lo[2] - p.z, p.z - S.aabb_hi[2]), 0.f);
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const int cls = (d2 >= PV_SORT_E0 * PV_SORT_E0) + (d2 >= PV_SORT_E1 * PV_SORT_E1) + (d2 >= PV_SORT_E2 * PV_SORT_E2);
    return cls * (PV_SORT_Q3_BINS * PV_SORT_Q5_BINS) + bin;
}

__device__ __forceinline__ void pv_cp_async4(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}
__device__ __forceinline__ void pv_cp_async16(void* smem, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}

// shared memory of the sorted kernel (dynamic: above the 48 KB static limit)
// stages of the key pass's load pipeline (one chunk each); they live in the staging area the main loop uses later
#ifndef PV_KEY_STAGES
#define PV_KEY_STAGES 6
#endif
struct PvSortSmem {
    static constexpr int ST = PV_ST;
    union {
        struct {
            float4 stA[2][PV_SB_THREADS], stB[2][PV_SB_THREADS];  // double-buffered staging of the NEXT iteration's configuration
            float st9[2][PV_SB_THREADS];                         // (SoA inputs; AoS inputs use stq)
            float stq[2][9][PV_SB_THREADS];
        };
        struct {  // key pass: q1..q4 and q6 of PV_KEY_STAGES chunks in flight
            float4 kA[PV_KEY_STAGES][PV_SB_THREADS];
            float k5[PV_KEY_STAGES][PV_SB_THREADS];
        };
    };
    unsigned short order[PV_ST];
    unsigned char key8[PV_ST];
    unsigned hist[PV_SORT_BUCKETS];
    unsigned vbits[PV_ST / 32];
    int cnt;
    // PV_COLD_SCENE: the configurations of the super-tile whose scene section is still owed are noted by the main loop
    // and worked off densely after it.  The list has an array of its own: the warps of the main loop run free (no
    // lockstep barrier any more, see the loop), so a fast warp may note configurations while a slow one still reads the
    // early entries of `order`.
    unsigned short owed[PV_ST];
    int n_owed;
};
// where the configurations come from: two float4 planes (+ optional ninth plane) or (n, 9) rows.  (The device-generated
// sweep keeps the unsorted kernel: it has no loads to hide, and parking the generated configurations in shared memory
// for the sort cost as much as the sort saved -- 8.87 vs 8.91 G checks/s, profiles/r1_notes.md.)
enum { PV_SRC_SOA = 0, PV_SRC_AOS = 1 };

template <int SRC, bool CARRY, bool YAW = false>
__global__ void __launch_bounds__(PV_SB_THREADS, PV_SB_MINB)
    pv_state_bits_sorted_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ qA,
                                const float4* __restrict__ qB, const float* __restrict__ q9,
                                const float* __restrict__ q_aos, int64_t n, uint32_t* __restrict__ bits,
                                const __grid_constant__ PvGather G) {
    constexpr bool AOS = SRC == PV_SRC_AOS;
    typedef PvSortSmem Smem;
    constexpr int ST_CHUNKS = Smem::ST / PV_SB_THREADS;
    extern __shared__ __align__(16) unsigned char pv_sort_smem_raw[];
    Smem& M = *reinterpret_cast<Smem*>(pv_sort_smem_raw);
    unsigned short* order = M.order;
    unsigned* hist = M.hist;
    unsigned* vbits = M.vbits;
    int& s_cnt = M.cnt;
    const int tid = threadIdx.x;
    const int64_t n_words = (n + 31) >> 5;
    const int64_t n_chunks = (n + PV_SB_THREADS - 1) / PV_SB_THREADS;
    // Programmatic dependent launch (see pv_launch_state_bits): the next state-check launch on this stream may place its
    // blocks on an SM as soon as this launch's block there has left -- its key pass, sort and checks then run in the
    // shadow of this launch's slower blocks instead of behind a grid-wide drain.  It reads configurations (which no state
    // kernel writes) and keeps its verdict words back until everything in front of it has completed (the wait below).
    asm volatile("griddepcontrol.launch_dependents;");
    if ((int64_t)blockIdx.x >= n_chunks) return;
    const int64_t my_chunks = (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    // local index L = jj * 512 + t of super-tile j0  <->  configuration (blockIdx.x + (j0 + jj) * gridDim.x) * 512 + t
#define PV_GI(jj, t) ((((int64_t)blockIdx.x + (j0 + (jj)) * (int64_t)gridDim.x) * PV_SB_THREADS) + (t))
    // 64-bit arithmetic once per super-tile (tile base pointers); inside a tile configurations are addressed by the
    // 32-bit offset PV_OFF from the tile's first configuration (at most 32 chunks x grid x 512 < 2^32)
#define PV_OFF(jj, t) ((unsigned)(jj) * stride + (unsigned)(t))
#define PV_Q03_OF(o)                                                                                               \
    (AOS ? make_float4(__ldg(t_aos + 9 * (o)), __ldg(t_aos + 9 * (o) + 1), __ldg(t_aos + 9 * (o) + 2), __ldg(t_aos + 9 * (o) + 3)) \
         : __ldg(tA + (o)))
    const unsigned stride = gridDim.x * PV_SB_THREADS;
    for (int64_t j0 = 0; j0 < my_chunks; j0 += ST_CHUNKS) {
        const int nc = (int)((my_chunks - j0 < (int64_t)ST_CHUNKS) ? (my_chunks - j0) : (int64_t)ST_CHUNKS);
        const int64_t base0 = PV_GI(0, 0);
        const unsigned n_rem = (unsigned)((n - base0 < (int64_t)0xffffffffll) ? (n - base0) : (int64_t)0xffffffffll);
        const float4* __restrict__ tA = AOS ? nullptr : qA + base0;
        const float4* __restrict__ tB = AOS ? nullptr : qB + base0;
        const float* __restrict__ t9 = (AOS || !q9) ? nullptr : q9 + base0;
        const float* __restrict__ t_aos = AOS ? q_aos + 9 * base0 : nullptr;
        if (tid < PV_SORT_BUCKETS) hist[tid] = 0;
        if (tid == 0) M.n_owed = 0;
        for (int w = tid; w < nc * (PV_SB_THREADS / 32); w += PV_SB_THREADS) vbits[w] = 0;
        __syncthreads();
        // pass 1: keys (kept as bytes for pass 2) and their histogram.  This is where the batch is first read from HBM, so
        // the loads run PV_KEY_STAGES - 1 chunks ahead of the key computation, as cp.async into shared memory (no
        // registers held): the ~1 us DRAM latency hides behind the ~60 instructions per key of the chunks in front (r2e
        // profile: with 4 register loads issued and then consumed the loop spent 267 of its 499 stall samples on
        // long_scoreboard, 13 % of the kernel).  A thread only reads back what it has copied itself: no barrier.
        {
            auto issue = [&](int j) {
                const unsigned o = PV_OFF(j, tid);
                if (j < nc && o < n_rem) {
                    const int st_ = j % PV_KEY_STAGES;
                    if constexpr (AOS) {
                        float* d = reinterpret_cast<float*>(&M.kA[st_][tid]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) pv_cp_async4(d + c, t_aos + 9 * o + c);
                        pv_cp_async4(&M.k5[st_][tid], t_aos + 9 * o + 5);
                    } else {
                        pv_cp_async16(&M.kA[st_][tid], tA + o);
                        pv_cp_async4(&M.k5[st_][tid], reinterpret_cast<const float*>(tB) + 4 * o + 1);
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");  // (an empty group keeps the count uniform)
            };
            for (int j = 0; j < PV_KEY_STAGES - 1; ++j) issue(j);
            for (int j = 0; j < nc; ++j) {
                issue(j + PV_KEY_STAGES - 1);
                asm volatile("cp.async.wait_group %0;" ::"n"(PV_KEY_STAGES - 1) : "memory");
                if (PV_OFF(j, tid) < n_rem) {
                    const float4 k = M.kA[j % PV_KEY_STAGES][tid];
                    const int key = pv_sort_key(k.x, k.y, k.z, k.w, M.k5[j % PV_KEY_STAGES][tid], S);
                    M.key8[j * PV_SB_THREADS + tid] = (unsigned char)key;
                    atomicAdd(&hist[key], 1u);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (tid < 32) {  // exclusive prefix sum: 8 buckets per lane + a warp scan
            constexpr int PER = PV_SORT_BUCKETS / 32;
            unsigned v[PER], sum = 0;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                v[j] = hist[tid * PER + j];
                sum += v[j];
            }
            unsigned incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
                if (tid >= d) incl += o;
            }
            unsigned run = incl - sum;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                hist[tid * PER + j] = run;
                run += v[j];
            }
            if (tid == 31) s_cnt = (int)incl;
        }
        __syncthreads();
        for (int jj = 0; jj < nc; ++jj) {  // pass 2: scatter the local indices (order inside a bucket does not matter)
            if (PV_OFF(jj, tid) < n_rem) order[atomicAdd(&hist[M.key8[jj * PV_SB_THREADS + tid]], 1u)] = (unsigned short)(jj * PV_SB_THREADS + tid);
        }
        __syncthreads();
        const int cnt = s_cnt;
        // the gathers of iteration r + 1 are issued (cp.async into shared memory, no registers held) before the check of
        // iteration r starts, so their L2 latency hides behind ~2000 instructions of work
        int L_next = 0;
#define PV_PREFETCH(r_)                                                                         \
    {                                                                                           \
        const int slot_ = (r_) * PV_SB_THREADS + tid;                                           \
        L_next = order[slot_ < cnt ? slot_ : cnt - 1];                                          \
        const unsigned i_ = PV_OFF(L_next / PV_SB_THREADS, L_next % PV_SB_THREADS);             \
        const int b_ = (r_) & 1;                                                                \
        if constexpr (AOS) {                                                                    \
            _Pragma("unroll") for (int j = 0; j < 9; ++j) pv_cp_async4(&M.stq[b_][j][tid], t_aos + 9 * i_ + j); \
        } else {                                                                                \
            pv_cp_async16(&M.stA[b_][tid], tA + i_);                                            \
            pv_cp_async16(&M.stB[b_][tid], tB + i_);                                            \
            if (q9) pv_cp_async4(&M.st9[b_][tid], t9 + i_);                                     \
        }                                                                                       \
        asm volatile("cp.async.commit_group;" ::: "memory");                                    \
    }
        PV_PREFETCH(0)
        for (int r = 0; r < nc; ++r) {
            const int slot = r * PV_SB_THREADS + tid;
            const bool in = slot < cnt;
            const int L = L_next;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            float q[9];
            if constexpr (AOS) {
#pragma unroll
                for (int j = 0; j < 9; ++j) q[j] = M.stq[r & 1][j][tid];
            } else {
                const float4 a = M.stA[r & 1][tid], b = M.stB[r & 1][tid];
                q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w;
                q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
                q[8] = q9 ? M.st9[r & 1][tid] : b.w;
            }
            if (r + 1 < nc) PV_PREFETCH(r + 1)
#if PV_SB_BAR_GROUPS == 0  // free-running warps (the default)
            __syncwarp();
#elif PV_SB_BAR_GROUPS > 1  // experiment: lockstep inside groups of warps only (named barriers)
            asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / (PV_SB_THREADS / PV_SB_BAR_GROUPS)), "n"(PV_SB_THREADS / PV_SB_BAR_GROUPS) : "memory");
#else
            __syncthreads();  // lockstep: the 16 warps of the block share instruction fetches
#endif
            PvAcc<PV_MODE_BITS> acc;
            constexpr bool COLD = PV_COLD_SCENE && PV_SB_SYNC < 3;
            const bool owes = pv_check_config<PV_MODE_BITS, true, PV_EXIT_NONE, PV_SB_SYNC, true, CARRY, true, COLD, YAW>(q, S, acc);
            if (in && !acc.hit) {
                // the few configurations that come near the scene boxes (~2 % in the goal scenes) are only NOTED here:
                // their scene section runs after the loop, densely packed, instead of in a sparsely populated warp now
                if (COLD && owes) M.owed[atomicAdd(&M.n_owed, 1)] = (unsigned short)L;
                else atomicOr(&vbits[L >> 5], 1u << (L & 31));
            }
        }
        __syncthreads();
        if constexpr (PV_COLD_SCENE && PV_SB_SYNC < 3) {
            const int n_owed = M.n_owed;
            // ~2 % of a super-tile owes: a few hundred configurations for 512 threads.  While there are lanes to spare a
            // configuration goes to 2 or 4 neighbouring lanes, each taking every 2nd / 4th box of the scene (the placement
            // is repeated per lane; the lanes were idle): the pass is the serial tail of the block, nothing overlaps it
            const int sh = (PV_OWED_SPLIT && n_owed * 4 <= PV_SB_THREADS) ? 2 : (PV_OWED_SPLIT && n_owed * 2 <= PV_SB_THREADS) ? 1 : 0;
            const int part = tid & ((1 << sh) - 1);
            for (int base = 0; base < n_owed; base += PV_SB_THREADS >> sh) {  // block-uniform trip count
                const int e = base + (tid >> sh);
                const bool have = e < n_owed;
                if (!__any_sync(0xffffffffu, have)) continue;  // (warp-uniform; no barrier inside this loop)
                const int L = M.owed[have ? e : n_owed - 1];
                const unsigned i_ = PV_OFF(L / PV_SB_THREADS, L % PV_SB_THREADS);
                float q[9];
                if constexpr (AOS) pv_load_aos(t_aos, (int64_t)i_, q);
                else pv_load_soa(tA, tB, t9, (int64_t)i_, q);
                PvPlaced P;
                pv_place<true, CARRY>(q, S, P);
                PvAcc<PV_MODE_BITS> acc;
                if (sh == 0) pv_scene_section<PV_MODE_BITS, true, PV_EXIT_NONE, 0, true, CARRY, YAW>(acc, P, S);  // (block-uniform)
                else pv_scene_section<PV_MODE_BITS, true, PV_EXIT_NONE, 0, true, CARRY, YAW>(acc, P, S, part, 1 << sh);
                // a configuration is valid when none of its lanes found a contact
                const unsigned hb = __ballot_sync(0xffffffffu, acc.hit);
                const unsigned mine = (hb >> ((tid & 31) & ~((1 << sh) - 1))) & ((1u << (1 << sh)) - 1u);
                if (have && part == 0 && !mine) atomicOr(&vbits[L >> 5], 1u << (L & 31));
            }
            __syncthreads();
        }
        // first global WRITE of this launch: not before every earlier operation on the stream has completed and is
        // visible (returns at once when the launch has no programmatic dependency, and after the first super-tile)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int wl = tid; wl < nc * (PV_SB_THREADS / 32); wl += PV_SB_THREADS) {
            const int64_t w = (PV_GI(wl / (PV_SB_THREADS / 32), 0) >> 5) + (wl % (PV_SB_THREADS / 32));
            if (w < n_words) pv_emit_word_thread(bits, G, w, vbits[wl]);
        }
        __syncthreads();  // the next super-tile clears hist / vbits
    }
#undef PV_PREFETCH
#undef PV_GI
#undef PV_Q03_OF
#undef PV_OFF
}

template <bool CULL, bool CARRY>
__global__ void __launch_bounds__(PV_THREADS, PV_MIN_BLOCKS)
    pv_state_margin_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ qA,
                           const float4* __restrict__ qB, const float* __restrict__ q9, int64_t n,
                           float* __restrict__ margin, int32_t* __restrict__ culprit) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float q[9];
        pv_load_soa(qA, qB, q9, i, q);
        PvAcc<PV_MODE_MARGIN> acc;
        pv_check_config<PV_MODE_MARGIN, CULL, PV_EXIT_NONE, 0, false, CARRY>(q, S, acc);
        margin[i] = acc.m;
        if (culprit) culprit[i] = acc.m < 0.f ? acc.code : 0;
    }
}

// diagnostics: every colliding pair of one state (the pair list robot.detect_collision() returns, planning.py:47-57)
template <bool CARRY>
__global__ void __launch_bounds__(PV_THREADS, PV_MIN_BLOCKS)
    pv_state_contacts_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ qA,
                             const float4* __restrict__ qB, const float* __restrict__ q9, int64_t n,
                             int32_t* __restrict__ codes, int32_t* __restrict__ count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float q[9];
        pv_load_soa(qA, qB, q9, i, q);
        PvAcc<PV_MODE_LIST> acc;
        pv_check_config<PV_MODE_LIST, false, PV_EXIT_NONE, 0, false, CARRY>(q, S, acc);
        count[i] = acc.n;
        for (int k = 0; k < PV_MAX_CONTACTS; ++k)
            codes[i * PV_MAX_CONTACTS + k] = (k < acc.n && k < PV_MAX_CONTACTS) ? acc.codes[k] : 0;
    }
}

// K1: forward kinematics only; [n][11][12] = position then row-major rotation per link.
template <bool FAST>
__global__ void __launch_bounds__(128)
    pv_fk_kernel(const __grid_constant__ PvScene S, const float4* __restrict__ qA, const float4* __restrict__ qB,
                 const float* __restrict__ q9, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float q[9];
        pv_load_soa(qA, qB, q9, i, q);
        float* o = out + i * 132;
        pv_fk_visit<FAST>(q, S.base[0], S.base[1], S.base[2], [&](auto lc, float3 p, float3 X, float3 Y, float3 Z) {
            constexpr int l = decltype(lc)::value;
            float* r = o + 12 * l;
            r[0] = p.x; r[1] = p.y; r[2] = p.z;
            r[3] = X.x; r[4] = Y.x; r[5] = Z.x;
            r[6] = X.y; r[7] = Y.y; r[8] = Z.y;
            r[9] = X.z; r[10] = Y.z; r[11] = Z.z;
        });
    }
}

// Config-5 sweep: configurations generated on device from a counter-based RNG, checked, bit-packed, counted.
template <bool CULL, bool CARRY>
__global__ void __launch_bounds__(PV_SB_THREADS, PV_SB_MINB)
    pv_sweep_kernel(const __grid_constant__ PvScene S, uint64_t first, int64_t n, unsigned seed, int fingers_open,
                    uint32_t* __restrict__ bits, unsigned long long* __restrict__ n_valid,
                    float* __restrict__ q_out, const __grid_constant__ PvGather G) {
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t n_words = (n + 31) >> 5;
    unsigned long long count = 0;
    for (int64_t wb = (int64_t)blockIdx.x * warps_per_block; wb < n_words; wb += (int64_t)gridDim.x * warps_per_block) {
        const int64_t w = wb + warp_in_block;
        const int64_t i = (w << 5) + lane;
        const bool in = i < n;
        float q[9];
        pv_sweep_config(first + (uint64_t)(in ? i : n - 1), seed, fingers_open != 0, q);
        if (q_out && in) {
#pragma unroll
            for (int j = 0; j < 9; ++j) q_out[9 * i + j] = q[j];
        }
        __syncthreads();  // lockstep: the warps of the block share instruction fetches
        PvAcc<PV_MODE_BITS> acc;
        pv_check_config<PV_MODE_BITS, CULL, PV_EXIT_NONE, 0, true, CARRY>(q, S, acc);
        const unsigned word = __ballot_sync(0xffffffffu, in && !acc.hit);
        if (w < n_words) {
            pv_emit_word(bits, G, w, word, lane);
            if (lane == 0) count += __popc(word);
        }
    }
    if (n_valid && lane == 0 && count) atomicAdd(n_valid, count);
}

// Sorted variant of the sweep (the default): the same super-tile sort as pv_state_bits_sorted_kernel, with the generated
// configurations parked in shared memory between the key pass and the check (generating them twice would cost ~210
// instructions per configuration, parking ~20).  A super-tile is PV_SWEEP_ST configurations (9 floats each: 144 KB).
// Super-tile sizes: as many configurations as fit the 227 KB of shared memory.  With the fingers fixed open (the
// BASELINE config-5 stream) q[7], q[8] are constants and only 7 floats per configuration are parked.
#ifndef PV_SWEEP_ST9
#define PV_SWEEP_ST9 5632  // 11 chunks, 9 floats each: 216 KB
#endif
#ifndef PV_SWEEP_ST7
#define PV_SWEEP_ST7 7168  // 14 chunks, 7 floats each: 219 KB
#endif
template <int NP, int ST_>
struct PvSweepSmem {
    static constexpr int ST = ST_;
    static_assert(ST_ % PV_SB_THREADS == 0 && ST_ <= 65536, "super-tile = whole chunks, indices fit 16 bits");
    float park[NP][ST_];
    unsigned short order[ST_];
    unsigned char key8[ST_];
    unsigned hist[PV_SORT_BUCKETS];
    unsigned vbits[ST_ / 32];
    int cnt;
    int n_owed;  // see PvSortSmem: the owed list reuses the front of `order`
};

template <bool CARRY, bool OPEN, bool YAW = false>
__global__ void __launch_bounds__(PV_SB_THREADS, PV_SB_MINB)
    pv_sweep_sorted_kernel(const __grid_constant__ PvScene S, uint64_t first, int64_t n, unsigned seed,
                           uint32_t* __restrict__ bits, unsigned long long* __restrict__ n_valid,
                           float* __restrict__ q_out, const __grid_constant__ PvGather G) {
    constexpr int NP = OPEN ? 7 : 9;
    typedef PvSweepSmem<NP, (OPEN ? PV_SWEEP_ST7 : PV_SWEEP_ST9)> Smem;
    constexpr int ST_CHUNKS = Smem::ST / PV_SB_THREADS;
    extern __shared__ __align__(16) unsigned char pv_sort_smem_raw[];
    Smem& M = *reinterpret_cast<Smem*>(pv_sort_smem_raw);
    const int tid = threadIdx.x;
    const int64_t n_words = (n + 31) >> 5;
    const int64_t n_chunks = (n + PV_SB_THREADS - 1) / PV_SB_THREADS;
    if ((int64_t)blockIdx.x >= n_chunks) return;
    const int64_t my_chunks = (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    unsigned long long count = 0;
#define PV_GI(jj, t) ((((int64_t)blockIdx.x + (j0 + (jj)) * (int64_t)gridDim.x) * PV_SB_THREADS) + (t))
    for (int64_t j0 = 0; j0 < my_chunks; j0 += ST_CHUNKS) {
        const int nc = (int)((my_chunks - j0 < (int64_t)ST_CHUNKS) ? (my_chunks - j0) : (int64_t)ST_CHUNKS);
        if (tid < PV_SORT_BUCKETS) M.hist[tid] = 0;
        if (tid == 0) M.n_owed = 0;
        for (int w = tid; w < nc * (PV_SB_THREADS / 32); w += PV_SB_THREADS) M.vbits[w] = 0;
        __syncthreads();
        for (int jj = 0; jj < nc; ++jj) {  // pass 1: generate, park, key, histogram
            const int64_t i = PV_GI(jj, tid);
            if (i < n) {
                float q[9];
                pv_sweep_config(first + (uint64_t)i, seed, OPEN, q);
                const int L = jj * PV_SB_THREADS + tid;
#pragma unroll
                for (int j = 0; j < NP; ++j) M.park[j][L] = q[j];
                if (q_out) {
#pragma unroll
                    for (int j = 0; j < 9; ++j) q_out[9 * i + j] = q[j];
                }
                const int key = pv_sort_key(q[0], q[1], q[2], q[3], q[5], S);
                M.key8[L] = (unsigned char)key;
                atomicAdd(&M.hist[key], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {  // exclusive prefix sum: 8 buckets per lane + a warp scan
            constexpr int PER = PV_SORT_BUCKETS / 32;
            unsigned v[PER], sum = 0;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                v[j] = M.hist[tid * PER + j];
                sum += v[j];
            }
            unsigned incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
                if (tid >= d) incl += o;
            }
            unsigned run = incl - sum;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                M.hist[tid * PER + j] = run;
                run += v[j];
            }
            if (tid == 31) M.cnt = (int)incl;
        }
        __syncthreads();
        for (int jj = 0; jj < nc; ++jj) {  // pass 2: scatter the local indices
            const int64_t i = PV_GI(jj, tid);
            const int L = jj * PV_SB_THREADS + tid;
            if (i < n) M.order[atomicAdd(&M.hist[M.key8[L]], 1u)] = (unsigned short)L;
        }
        __syncthreads();
        const int cnt = M.cnt;
        for (int r = 0; r < nc; ++r) {
            const int slot = r * PV_SB_THREADS + tid;
            const bool in = slot < cnt;
            const int L = M.order[in ? slot : cnt - 1];
            float q[9];
#pragma unroll
            for (int j = 0; j < NP; ++j) q[j] = M.park[j][L];
            if (OPEN) q[7] = q[8] = 0.04f;  // what pv_sweep_config sets
            __syncthreads();  // lockstep: the 16 warps of the block share instruction fetches
            PvAcc<PV_MODE_BITS> acc;
            const bool owes = pv_check_config<PV_MODE_BITS, true, PV_EXIT_NONE, 0, true, CARRY, true, (PV_COLD_SCENE != 0), YAW>(q, S, acc);
            if (in && !acc.hit) {
                if (PV_COLD_SCENE && owes) M.order[atomicAdd(&M.n_owed, 1)] = (unsigned short)L;
                else atomicOr(&M.vbits[L >> 5], 1u << (L & 31));
            }
        }
        __syncthreads();
        if constexpr (PV_COLD_SCENE != 0) {  // the scene sections still owed, densely packed (see the state kernel)
            const int n_owed = M.n_owed;
            // (lanes to spare: a configuration's boxes dealt out to 2 or 4 neighbouring lanes, see the state kernel)
            const int sh = (PV_OWED_SPLIT && n_owed * 4 <= PV_SB_THREADS) ? 2 : (PV_OWED_SPLIT && n_owed * 2 <= PV_SB_THREADS) ? 1 : 0;
            const int part = tid & ((1 << sh) - 1);
            for (int base = 0; base < n_owed; base += PV_SB_THREADS >> sh) {
                const int e = base + (tid >> sh);
                const bool have = e < n_owed;
                if (!__any_sync(0xffffffffu, have)) continue;  // (warp-uniform; no barrier inside this loop)
                const int L = M.order[have ? e : n_owed - 1];
                float q[9];
#pragma unroll
                for (int j = 0; j < NP; ++j) q[j] = M.park[j][L];
                if (OPEN) q[7] = q[8] = 0.04f;
                PvPlaced P;
                pv_place<true, CARRY>(q, S, P);
                PvAcc<PV_MODE_BITS> acc;
                if (sh == 0) pv_scene_section<PV_MODE_BITS, true, PV_EXIT_NONE, 0, true, CARRY, YAW>(acc, P, S);
                else pv_scene_section<PV_MODE_BITS, true, PV_EXIT_NONE, 0, true, CARRY, YAW>(acc, P, S, part, 1 << sh);
                const unsigned hb = __ballot_sync(0xffffffffu, acc.hit);
                const unsigned mine = (hb >> ((tid & 31) & ~((1 << sh) - 1))) & ((1u << (1 << sh)) - 1u);
                if (have && part == 0 && !mine) atomicOr(&M.vbits[L >> 5], 1u << (L & 31));
            }
            __syncthreads();
        }
        for (int wl = tid; wl < nc * (PV_SB_THREADS / 32); wl += PV_SB_THREADS) {
            const int64_t w = (PV_GI(wl / (PV_SB_THREADS / 32), 0) >> 5) + (wl % (PV_SB_THREADS / 32));
            if (w < n_words) {
                const unsigned word = M.vbits[wl];
                pv_emit_word_thread(bits, G, w, word);
                count += __popc(word);
            }
        }
        __syncthreads();  // the next super-tile clears hist / vbits
    }
#undef PV_GI
    if (n_valid) {  // one atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) count += __shfl_down_sync(0xffffffffu, count, d);
        if ((tid & 31) == 0 && count) atomicAdd(n_valid, count);
    }
}

// FP32 issue-rate probe: 8 independent FFMA chains per thread.
__global__ void __launch_bounds__(256) pv_fp32_peak_kernel(int iters, float seed, float* out) {
    float a0 = seed, a1 = seed + 1.f, a2 = seed + 2.f, a3 = seed + 3.f;
    float a4 = seed + 4.f, a5 = seed + 5.f, a6 = seed + 6.f, a7 = seed + 7.f;
    const float m = 0.999f + 1e-6f * (float)threadIdx.x, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) out[0] = r;  // never true in practice; keeps the chains alive
}

// =========================================================================================================
// host side
// =========================================================================================================
static int pv_check_handle(const PvHandle* h) { return (h && h->magic == PV_HANDLE_MAGIC) ? 0 : PV_ERR_BAD_HANDLE; }

int pv_grid_for(PvHandle* h, const void* kernel, int threads, int64_t warps_needed) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 1;
    int64_t blocks = (int64_t)h->sm_count * occ;
    const int64_t need = (warps_needed * 32 + threads - 1) / threads;
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" {

const char* pv_version(void) { return "panda_validity 0.1 (sm_100a)"; }

const char* pv_last_error(const PvHandle* h) { return (h && h->magic == PV_HANDLE_MAGIC) ? h->err : g_create_error; }

long long pv_launch_count(const PvHandle* h) { return pv_check_handle(h) ? -1 : h->launches; }

int pv_model_info(int* n_spheres, int* n_boxes, int* n_ss_pairs, int* n_sb_pairs) {
    if (n_spheres) *n_spheres = PV_N_SPHERES;
    if (n_boxes) *n_boxes = PV_N_BOXES;
    if (n_ss_pairs) *n_ss_pairs = PV_N_SS_PAIRS;
    if (n_sb_pairs) *n_sb_pairs = PV_N_SB_PAIRS;
    return PV_OK;
}

int pv_joint_limits(float lower[9], float upper[9]) {
    const float lo[9] = PV_Q_LOWER, hi[9] = PV_Q_UPPER;
    if (!lower || !upper) return PV_ERR_BAD_ARG;
    memcpy(lower, lo, sizeof(lo));
    memcpy(upper, hi, sizeof(hi));
    return PV_OK;
}

int pv_create(int device, PvHandle** out) {
    if (!out) return PV_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        snprintf(g_create_error, sizeof(g_create_error), "no CUDA device visible (%s); there is no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PV_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        snprintf(g_create_error, sizeof(g_create_error), "device %d out of range (0..%d)", device, count - 1);
        return PV_ERR_BAD_ARG;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        snprintf(g_create_error, sizeof(g_create_error), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return PV_ERR_CUDA;
    }
    if (prop.major != 10) {
        snprintf(g_create_error, sizeof(g_create_error),
                 "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return PV_ERR_NO_DEVICE;
    }
    PvHandle* h = new PvHandle();
    memset(h, 0, sizeof(*h));
    h->magic = PV_HANDLE_MAGIC;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->scene.attached = -1;
    h->scene.flags = PV_FLAG_SELF;
    h->scene.base[2] = 0.01f;
    h->cull = 2;
    h->launch_overlap = 1;
    {
        const char* e2 = getenv("PV_EDGE_CERT2");  // developer switch for same-box A/B runs
        h->edge_cert2 = !(e2 && e2[0] == '0');
    }
    PvDeviceGuard guard(device);  // the streams are created on `device`; the caller's current device is put back
    if ((e = cudaGetLastError()) != cudaSuccess) {
        snprintf(g_create_error, sizeof(g_create_error), "cudaSetDevice: %s", cudaGetErrorString(e));
        delete h;
        return PV_ERR_CUDA;
    }
    for (int i = 0; i < PV_N_STREAMS; ++i) {
        if ((e = cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking)) != cudaSuccess) {
            snprintf(g_create_error, sizeof(g_create_error), "cudaStreamCreate: %s", cudaGetErrorString(e));
            delete h;
            return PV_ERR_CUDA;
        }
    }
    *out = h;
    return PV_OK;
}

void pv_destroy(PvHandle* h) {
    if (pv_check_handle(h)) return;
    PvDeviceGuard guard(h->device);
    for (int i = 0; i < PV_N_STREAMS; ++i) {
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
        if (h->stage_q[i]) cudaFree(h->stage_q[i]);
        if (h->stage_q2[i]) cudaFree(h->stage_q2[i]);
        if (h->stage_bits[i]) cudaFree(h->stage_bits[i]);
    }
    if (h->rrtc_buf) cudaFree(h->rrtc_buf);
    if (h->rrtc_host) cudaFreeHost(h->rrtc_host);
    if (h->rrtc_rows_host) cudaFreeHost(h->rrtc_rows_host);
    if (h->plan_host) cudaFreeHost(h->plan_host);
    if (h->small_host) cudaFreeHost(h->small_host);
    if (h->ik_buf) cudaFree(h->ik_buf);
    if (h->pool) cudaMemPoolDestroy(h->pool);
    h->magic = 0;
    delete h;
}

int pv_set_scene(PvHandle* h, const float* h_obb, int n_obb, float table_z, const float base_xyz[3]) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    if (n_obb < 0 || n_obb > PV_MAX_OBB || (n_obb > 0 && !h_obb)) {
        snprintf(h->err, sizeof(h->err), "pv_set_scene: n_obb=%d outside 0..%d (or null buffer)", n_obb, PV_MAX_OBB);
        return PV_ERR_BAD_ARG;
    }
    // build into a copy and commit at the end: a rejected call leaves the previous scene intact
    PvScene S = h->scene;
    memset(S.obb, 0, sizeof(S.obb));
    S.yaw_only_mask = 0;
    for (int b = 0; b < n_obb; ++b) {
        const float* o = h_obb + 16 * b;
        for (int k = 0; k < 15; ++k) {
            if (!isfinite(o[k])) {
                snprintf(h->err, sizeof(h->err), "pv_set_scene: box %d has a non-finite entry", b);
                return PV_ERR_BAD_ARG;
            }
            S.obb[b][k] = o[k];
        }
        if (!(o[3] > 0.f && o[4] > 0.f && o[5] > 0.f)) {
            snprintf(h->err, sizeof(h->err), "pv_set_scene: box %d has non-positive half extents", b);
            return PV_ERR_BAD_ARG;
        }
        S.obb[b][15] = sqrtf(o[3] * o[3] + o[4] * o[4] + o[5] * o[5]);
        const float* R = o + 6;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                const float dot = R[3 * i] * R[3 * j] + R[3 * i + 1] * R[3 * j + 1] + R[3 * i + 2] * R[3 * j + 2];
                if (fabsf(dot - (i == j ? 1.f : 0.f)) > 1e-3f) {
                    snprintf(h->err, sizeof(h->err), "pv_set_scene: box %d rotation is not orthonormal", b);
                    return PV_ERR_BAD_ARG;
                }
            }
        if (R[2] == 0.f && R[5] == 0.f && R[6] == 0.f && R[7] == 0.f && R[8] == 1.f) S.yaw_only_mask |= 1u << b;
    }
    S.n_obb = n_obb;
    S.table_z = table_z;
    if (base_xyz) {
        S.base[0] = base_xyz[0];
        S.base[1] = base_xyz[1];
        S.base[2] = base_xyz[2];
    }
    if (S.attached >= n_obb) {
        S.attached = -1;
        S.carry = 0;
    }
    // static reach masks: which link groups / gripper boxes can touch which scene box at all
    {
        const float link_reach[8] = PV_LINK_REACH, box_reach[3] = PV_BOX_REACH;
        const float slack = 1e-3f;
        const float sx = S.base[0], sy = S.base[1], sz = S.base[2] + 0.333f;
        for (int b = 0; b < PV_MAX_OBB; ++b) S.reach_mask[b] = 0;
        for (int b = 0; b < n_obb; ++b) {
            const float* o = S.obb[b];
            const float dx = o[0] - sx, dy = o[1] - sy, dz = o[2] - sz;
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz) - o[15];
            unsigned m = 0;
            for (int l = 1; l < 8; ++l)
                if (dist < link_reach[l] + slack) m |= 1u << l;
            for (int k = 0; k < 3; ++k)
                if (dist < box_reach[k] + slack) m |= 1u << (8 + k);
#define PV_L0_REACH(i, link, cx, cy, cz, r)                                                         \
    {                                                                                               \
        const float ex = o[0] - (S.base[0] + cx), ey = o[1] - (S.base[1] + cy), ez = o[2] - (S.base[2] + cz); \
        if (sqrtf(ex * ex + ey * ey + ez * ez) < (r) + o[15] + slack) m |= 1u;                      \
    }
            PV_SPHERES_LINK0(PV_L0_REACH)
#undef PV_L0_REACH
            S.reach_mask[b] = (unsigned short)m;
        }
    }
    // scene-level cull: bounds of all boxes, padded per link group (pv_check_config skips the box loop for a
    // configuration none of whose group balls enters them)
    {
        unsigned any = 0;
        for (int k = 0; k < 3; ++k) {
            S.aabb_lo[k] = 1e30f;
            S.aabb_hi[k] = -1e30f;
        }
        for (int b = 0; b < n_obb; ++b) {
            const float* o = S.obb[b];
            any |= S.reach_mask[b];
            for (int k = 0; k < 3; ++k) {  // world axis k: extent = sum_j |R[k][j]| h[j]
                const float e = fabsf(o[6 + 3 * k]) * o[3] + fabsf(o[7 + 3 * k]) * o[4] + fabsf(o[8 + 3 * k]) * o[5];
                S.aabb_lo[k] = fminf(S.aabb_lo[k], o[k] - e - 1e-6f);
                S.aabb_hi[k] = fmaxf(S.aabb_hi[k], o[k] + e + 1e-6f);
            }
        }
        S.group_any = (any & 0xffu) | ((any & 0x700u) ? 0x100u : 0u);
        memset(S.gpad, 0, sizeof(S.gpad));
#define PV_GPAD(l, cs, br)                                              \
    for (int k = 0; k < 3; ++k) {                                       \
        S.gpad[l][k] = S.aabb_lo[k] - ((br) + 2.0f * PV_CULL_SLACK);    \
        S.gpad[l][3 + k] = S.aabb_hi[k] + ((br) + 2.0f * PV_CULL_SLACK); \
    }
        PV_LINK_GROUPS(PV_GPAD)
#undef PV_GPAD
    }
    h->scene = S;
    h->has_scene = 1;
    h->all_yaw = S.yaw_only_mask == (S.n_obb >= 32 ? 0xffffffffu : ((1u << S.n_obb) - 1u));
    return PV_OK;
}

int pv_set_attached(PvHandle* h, int obb_index) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    if (obb_index < -1 || obb_index >= h->scene.n_obb) {
        snprintf(h->err, sizeof(h->err), "pv_set_attached: index %d outside -1..%d", obb_index, h->scene.n_obb - 1);
        return PV_ERR_BAD_ARG;
    }
    h->scene.attached = obb_index;
    h->scene.carry = 0;
    return PV_OK;
}

int pv_set_carried(PvHandle* h, int obb_index, const float* hand_from_box, float contact_allowance) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    if (obb_index == -1) {
        h->scene.attached = -1;
        h->scene.carry = 0;
        return PV_OK;
    }
    if (obb_index < 0 || obb_index >= h->scene.n_obb || !hand_from_box) {
        snprintf(h->err, sizeof(h->err), "pv_set_carried: index %d outside -1..%d (or null pose)", obb_index,
                 h->scene.n_obb - 1);
        return PV_ERR_BAD_ARG;
    }
    const float* R = hand_from_box;
    for (int k = 0; k < 12; ++k)
        if (!isfinite(hand_from_box[k])) {
            snprintf(h->err, sizeof(h->err), "pv_set_carried: non-finite pose entry");
            return PV_ERR_BAD_ARG;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            const float dot = R[3 * i] * R[3 * j] + R[3 * i + 1] * R[3 * j + 1] + R[3 * i + 2] * R[3 * j + 2];
            if (fabsf(dot - (i == j ? 1.f : 0.f)) > 1e-3f) {
                snprintf(h->err, sizeof(h->err), "pv_set_carried: rotation is not orthonormal");
                return PV_ERR_BAD_ARG;
            }
        }
    const float* o = h->scene.obb[obb_index];
    if (!(contact_allowance >= 0.f) || !(contact_allowance < fminf(o[3], fminf(o[4], o[5])))) {
        snprintf(h->err, sizeof(h->err), "pv_set_carried: contact allowance %g outside [0, smallest half extent)",
                 (double)contact_allowance);
        return PV_ERR_BAD_ARG;
    }
    for (int k = 0; k < 3; ++k) h->scene.carry_h[k] = o[3 + k] - contact_allowance;
    h->scene.carry_br = sqrtf(h->scene.carry_h[0] * h->scene.carry_h[0] + h->scene.carry_h[1] * h->scene.carry_h[1] +
                              h->scene.carry_h[2] * h->scene.carry_h[2]);
    for (int k = 0; k < 9; ++k) h->scene.carry_R[k] = R[k];
    for (int k = 0; k < 3; ++k) h->scene.carry_t[k] = hand_from_box[9 + k];
    h->scene.attached = obb_index;
    h->scene.carry = 1;
    return PV_OK;
}

int pv_set_gather(PvHandle* h, const void* d_peer_ptrs, int n_peers, void* d_multicast, long long word_offset,
                  long long word_capacity) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    if (n_peers < 0 || n_peers > 32 || word_offset < 0 || word_capacity < 0 || (n_peers > 0 && !d_peer_ptrs && !d_multicast)) {
        snprintf(h->err, sizeof(h->err), "pv_set_gather: bad arguments");
        return PV_ERR_BAD_ARG;
    }
    memset(&h->gather, 0, sizeof(h->gather));
    if (n_peers > 0) {
        h->gather.peers = (uint32_t* const*)d_peer_ptrs;
        h->gather.mc = (uint32_t*)d_multicast;
        h->gather.n_peers = n_peers;
        h->gather.word_off = word_offset;
        h->gather.word_cap = word_capacity;
    }
    return PV_OK;
}

int pv_set_flags(PvHandle* h, unsigned flags) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    h->scene.flags = flags & (PV_FLAG_SELF | PV_FLAG_LIMITS);
    return PV_OK;
}

// test / profiling hook for the state kernel: 0 = brute force (every kept pair tested), 1 = per-lane
// bounding-ball culling in front of each block of tests (default).  Verdicts are bit-identical.
int pv_set_culling(PvHandle* h, int on) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    h->cull = on < 0 ? 0 : (on > 2 ? 2 : on);  // 0 brute force, 1 per-lane culling, 2 tile-sorted + culling
    return PV_OK;
}

// 1 (default): consecutive pv_check_states launches on one stream overlap (programmatic dependent launch); 0: every
// launch waits for the previous one to drain, as plain stream order would have it.  Same verdict words either way.
int pv_set_launch_overlap(PvHandle* h, int on) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    h->launch_overlap = on ? 1 : 0;
    return PV_OK;
}

#define PV_PRECHECK(h, n)                                                   \
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;                       \
    if (!(h)->has_scene) {                                                  \
        snprintf((h)->err, sizeof((h)->err), "no scene set (pv_set_scene)"); \
        return PV_ERR_NO_SCENE;                                             \
    }                                                                       \
    if ((n) < 0) {                                                          \
        snprintf((h)->err, sizeof((h)->err), "negative count");             \
        return PV_ERR_BAD_ARG;                                              \
    }                                                                       \
    if ((n) == 0) return PV_OK;                                             \
    PvDeviceGuard pv_guard_((h)->device);

static int pv_fk_impl(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9, int64_t n, float* d_pose_out,
                      void* stream, bool verdict_path) {
    PV_PRECHECK(h, n);
    if (!d_qA || !d_qB || !d_pose_out) return PV_ERR_BAD_ARG;
    const int threads = 128;
    int64_t blocks = (n + threads - 1) / threads;
    if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
    if (verdict_path && PV_FAST_TRIG)
        pv_fk_kernel<true><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(h->scene, (const float4*)d_qA,
                                                                              (const float4*)d_qB, d_q9, n, d_pose_out);
    else
        pv_fk_kernel<false><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(h->scene, (const float4*)d_qA,
                                                                               (const float4*)d_qB, d_q9, n, d_pose_out);
    h->launches++;
    PV_CUDA(h, cudaGetLastError());
    return PV_OK;
}

int pv_fk(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9, int64_t n, float* d_pose_out,
          void* stream) {
    return pv_fk_impl(h, d_qA, d_qB, d_q9, n, d_pose_out, stream, false);
}

int pv_fk_verdict_path(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9, int64_t n,
                       float* d_pose_out, void* stream) {
    return pv_fk_impl(h, d_qA, d_qB, d_q9, n, d_pose_out, stream, true);
}

int pv_state_contacts(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9, int64_t n,
                      int32_t* d_codes, int32_t* d_count, void* stream) {
    PV_PRECHECK(h, n);
    if (!d_qA || !d_qB || !d_codes || !d_count) return PV_ERR_BAD_ARG;
    if (h->scene.carry) {
        int grid = pv_grid_for(h, (const void*)pv_state_contacts_kernel<true>, PV_THREADS, (n + 31) / 32);
        pv_state_contacts_kernel<true><<<grid, PV_THREADS, 0, (cudaStream_t)stream>>>(
            h->scene, (const float4*)d_qA, (const float4*)d_qB, d_q9, n, d_codes, d_count);
    } else {
        int grid = pv_grid_for(h, (const void*)pv_state_contacts_kernel<false>, PV_THREADS, (n + 31) / 32);
        pv_state_contacts_kernel<false><<<grid, PV_THREADS, 0, (cudaStream_t)stream>>>(
            h->scene, (const float4*)d_qA, (const float4*)d_qB, d_q9, n, d_codes, d_count);
    }
    h->launches++;
    PV_CUDA(h, cudaGetLastError());
    return PV_OK;
}

static int pv_launch_state_bits(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9,
                                const float* d_aos, int64_t n, uint32_t* d_bits, cudaStream_t st,
                                bool force_unsorted = false, bool host_call = false) {
    // the fused gather forwards the words of DEVICE-buffer calls only: a host-buffer call (whatever staging layout it uses
    // internally) numbers its words per chunk and must never write into the ranks' gather buffers
    const PvGather gather_ = (d_aos || host_call) ? PvGather{} : h->gather;
    const int64_t words = (n + 31) / 32;
#define PV_LAUNCH_SB(AOS, CULL, CARRY)                                                                        \
    {                                                                                                         \
        int grid = pv_grid_for(h, (const void*)pv_state_bits_kernel<AOS, CULL, CARRY>, PV_SB_THREADS, words); \
        pv_state_bits_kernel<AOS, CULL, CARRY><<<grid, PV_SB_THREADS, 0, st>>>(                               \
            h->scene, (const float4*)d_qA, (const float4*)d_qB, d_q9, d_aos, n, d_bits,                       \
            gather_);                                                                  \
    }
#define PV_LAUNCH_SORTED(AOS_, CARRY, YAW_)                                                                   \
    {                                                                                                         \
        int64_t chunks = (n + PV_SB_THREADS - 1) / PV_SB_THREADS;                                             \
        int grid = (int)(chunks < (int64_t)h->sm_count ? chunks : (int64_t)h->sm_count);                      \
        constexpr int SRC_ = (AOS_) ? PV_SRC_AOS : PV_SRC_SOA;                                                \
        const unsigned bit_ = 1u << (2 * SRC_ + (CARRY ? 1 : 0) + (YAW_ ? 16 : 0));                           \
        if (!(h->smem_attr_mask & bit_)) {                                                                    \
            PV_CUDA(h, cudaFuncSetAttribute(pv_state_bits_sorted_kernel<SRC_, CARRY, YAW_>,                   \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PvSortSmem))); \
            h->smem_attr_mask |= bit_;                                                                        \
        }                                                                                                     \
        /* launched with programmatic stream serialization: consecutive state-check launches overlap (kernel header) */ \
        cudaLaunchConfig_t cfg_ = {};                                                                         \
        cfg_.gridDim = dim3((unsigned)grid);                                                                  \
        cfg_.blockDim = dim3(PV_SB_THREADS);                                                                  \
        cfg_.dynamicSmemBytes = sizeof(PvSortSmem);                                                           \
        cfg_.stream = st;                                                                                     \
        cudaLaunchAttribute at_[1];                                                                           \
        at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                       \
        at_[0].val.programmaticStreamSerializationAllowed = 1;                                                \
        cfg_.attrs = at_;                                                                                     \
        cfg_.numAttrs = h->launch_overlap ? 1 : 0;                                                            \
        PV_CUDA(h, cudaLaunchKernelEx(&cfg_, pv_state_bits_sorted_kernel<SRC_, CARRY, YAW_>, h->scene, (const float4*)d_qA, \
                                      (const float4*)d_qB, d_q9, d_aos, n, d_bits, gather_)); \
    }
    if (h->cull == 2 && !force_unsorted) {  // tile-sorted + per-lane culling (the default)
        if (h->scene.carry) {
            if (d_aos) PV_LAUNCH_SORTED(true, true, false) else PV_LAUNCH_SORTED(false, true, false)
        } else if (h->all_yaw) {
            if (d_aos) PV_LAUNCH_SORTED(true, false, true) else PV_LAUNCH_SORTED(false, false, true)
        } else {
            if (d_aos) PV_LAUNCH_SORTED(true, false, false) else PV_LAUNCH_SORTED(false, false, false)
        }
    } else if (h->scene.carry) {  // carry mode always culls (the brute-force variant exists for the A/B identity test)
        if (d_aos) PV_LAUNCH_SB(true, true, true) else PV_LAUNCH_SB(false, true, true)
    } else if (d_aos) {
        if (h->cull) PV_LAUNCH_SB(true, true, false) else PV_LAUNCH_SB(true, false, false)
    } else {
        if (h->cull) PV_LAUNCH_SB(false, true, false) else PV_LAUNCH_SB(false, false, false)
    }
#undef PV_LAUNCH_SB
#undef PV_LAUNCH_SORTED
    h->launches++;
    PV_CUDA(h, cudaGetLastError());
    return PV_OK;
}

int pv_check_states(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9, int64_t n,
                    uint32_t* d_bits, void* stream) {
    PV_PRECHECK(h, n);
    if (!d_qA || !d_qB || !d_bits) return PV_ERR_BAD_ARG;
    return pv_launch_state_bits(h, d_qA, d_qB, d_q9, nullptr, n, d_bits, (cudaStream_t)stream);
}

int pv_state_margins(PvHandle* h, const float* d_qA, const float* d_qB, const float* d_q9, int64_t n,
                     float* d_margin, int32_t* d_culprit, void* stream) {
    PV_PRECHECK(h, n);
    if (!d_qA || !d_qB || !d_margin) return PV_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        // margins are a diagnostic: always brute force (culling only guarantees that contacts are never
        // missed, it may skip the far-away primitive that defines a positive clearance)
        if (h->scene.carry) {
            int grid = pv_grid_for(h, (const void*)pv_state_margin_kernel<false, true>, PV_THREADS, (n + 31) / 32);
            pv_state_margin_kernel<false, true><<<grid, PV_THREADS, 0, st>>>(
                h->scene, (const float4*)d_qA, (const float4*)d_qB, d_q9, n, d_margin, d_culprit);
        } else {
            int grid = pv_grid_for(h, (const void*)pv_state_margin_kernel<false, false>, PV_THREADS, (n + 31) / 32);
            pv_state_margin_kernel<false, false><<<grid, PV_THREADS, 0, st>>>(
                h->scene, (const float4*)d_qA, (const float4*)d_qB, d_q9, n, d_margin, d_culprit);
        }
    }
    h->launches++;
    PV_CUDA(h, cudaGetLastError());
    return PV_OK;
}

// ---- host-buffer entry points: chunked H2D -> kernel -> D2H pipeline over PV_N_STREAMS streams -----------
static int pv_ensure_stage(PvHandle* h) {
    for (int i = 0; i < PV_N_STREAMS; ++i) {
        if (!h->stage_q[i]) PV_CUDA(h, cudaMalloc(&h->stage_q[i], (size_t)PV_HOST_CHUNK * 9 * sizeof(float)));
        if (!h->stage_q2[i]) PV_CUDA(h, cudaMalloc(&h->stage_q2[i], (size_t)PV_HOST_CHUNK * 9 * sizeof(float)));
        if (!h->stage_bits[i]) PV_CUDA(h, cudaMalloc(&h->stage_bits[i], (size_t)(PV_HOST_CHUNK / 32) * sizeof(uint32_t)));
    }
    return PV_OK;
}

// Chunk schedule of the pipelined state calls: full chunks, then a SHORT last one.  The copies run back to back on the one
// host-to-device engine, so the call ends one kernel + one D2H after the last copy has landed: that kernel should be small.
#ifndef PV_HOST_TAIL
#define PV_HOST_TAIL 32768  // configurations in the last chunk of a multi-chunk call (multiple of 32; 0: equal chunks)
#endif
static inline int64_t pv_host_chunk(int64_t left) {
    if (left > (int64_t)PV_HOST_CHUNK + PV_HOST_TAIL) return PV_HOST_CHUNK;
    if (PV_HOST_TAIL > 0 && left > 2 * (int64_t)PV_HOST_TAIL) return (left - PV_HOST_TAIL) & ~(int64_t)31;
    return left < PV_HOST_CHUNK ? left : PV_HOST_CHUNK;
}

// Small host batches -- the reference's callback shape is ONE state per call (planning.py:209-219), a plan's waypoint
// validation a few hundred motions -- go through host-mapped pinned memory: the kernel reads the rows and writes the verdict
// words over PCIe itself, so a call is one launch and one synchronisation instead of copy + launch + copy (34 -> ~12 us
// for one state).  No sort for these sizes: the unsorted kernel returns the same words.
#ifndef PV_HOST_SMALL
#define PV_HOST_SMALL 2048
#endif
static int pv_ensure_small(PvHandle* h) {
    if (!h->small_host) {
        const size_t bytes = (size_t)PV_HOST_SMALL * 9 * sizeof(float) * 2 + (size_t)PV_HOST_SMALL * sizeof(uint32_t);
        PV_CUDA(h, cudaHostAlloc(&h->small_host, bytes, cudaHostAllocMapped));
    }
    return PV_OK;
}

int pv_check_states_host(PvHandle* h, const float* h_q, int64_t n, uint32_t* h_bits) {
    PV_PRECHECK(h, n);
    if (!h_q || !h_bits) return PV_ERR_BAD_ARG;
    int rc;
    if (n <= PV_HOST_SMALL) {
        if ((rc = pv_ensure_small(h))) return rc;
        float* mq = (float*)h->small_host;
        uint32_t* mb = (uint32_t*)(mq + (size_t)PV_HOST_SMALL * 18);
        memcpy(mq, h_q, (size_t)n * 9 * sizeof(float));
        if ((rc = pv_launch_state_bits(h, nullptr, nullptr, nullptr, mq, n, mb, h->streams[0], true))) return rc;
        PV_CUDA(h, cudaStreamSynchronize(h->streams[0]));
        memcpy(h_bits, mb, (size_t)((n + 31) / 32) * sizeof(uint32_t));
        return PV_OK;
    }
    rc = pv_ensure_stage(h);
    if (rc) return rc;
    // chunk size: whole input for small batches, else PV_HOST_CHUNK (multiple of 32) so copies overlap compute
    int64_t done = 0;
    int slot = 0;
    while (done < n) {
        const int64_t m = pv_host_chunk(n - done);
        cudaStream_t st = h->streams[slot];
        PV_CUDA(h, cudaMemcpyAsync(h->stage_q[slot], h_q + done * 9, (size_t)m * 9 * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
        rc = pv_launch_state_bits(h, nullptr, nullptr, nullptr, h->stage_q[slot], m, h->stage_bits[slot], st);
        if (rc) return rc;
        PV_CUDA(h, cudaMemcpyAsync(h_bits + done / 32, h->stage_bits[slot], (size_t)((m + 31) / 32) * sizeof(uint32_t),
                                   cudaMemcpyDeviceToHost, st));
        done += m;
        slot = (slot + 1) % PV_N_STREAMS;
    }
    for (int i = 0; i < PV_N_STREAMS; ++i) PV_CUDA(h, cudaStreamSynchronize(h->streams[i]));
    return PV_OK;
}

// rows of 7 arm joint values -> the SoA planes the validity kernels read (q8 = finger_left in B.w, q9 = finger_right)
__global__ void __launch_bounds__(256) pv_arm_rows_to_planes_kernel(const float* __restrict__ rows, int64_t n, float f_left,
                                                                    float f_right, float4* __restrict__ A,
                                                                    float4* __restrict__ B, float* __restrict__ q9) {
    // a block stages its 256 rows (1 792 consecutive floats) through shared memory: coalesced reads, float4 writes
    __shared__ float s[256 * 7];
    const int64_t base = (int64_t)blockIdx.x * 256;
    const int here = (int)(n - base < 256 ? n - base : 256);
    for (int k = threadIdx.x; k < here * 7; k += 256) s[k] = __ldg(rows + base * 7 + k);
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= here) return;
    const float* r = s + t * 7;
    A[base + t] = make_float4(r[0], r[1], r[2], r[3]);
    B[base + t] = make_float4(r[4], r[5], r[6], f_left);
    if (q9) q9[base + t] = f_right;
}

int pv_check_states_host_arm(PvHandle* h, const float* h_q7, int64_t n, float finger_left, float finger_right,
                             uint32_t* h_bits) {
    PV_PRECHECK(h, n);
    if (!h_q7 || !h_bits) return PV_ERR_BAD_ARG;
    int rc = pv_ensure_stage(h);
    if (rc) return rc;
    int64_t done = 0;
    int slot = 0;
    while (done < n) {
        const int64_t m = pv_host_chunk(n - done);
        cudaStream_t st = h->streams[slot];
        PV_CUDA(h, cudaMemcpyAsync(h->stage_q[slot], h_q7 + done * 7, (size_t)m * 7 * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
        // the second staging buffer (36 B per configuration) holds the planes: A | B | q9
        float* pA = h->stage_q2[slot];
        float* pB = pA + 4 * (size_t)PV_HOST_CHUNK;
        float* p9 = finger_left == finger_right ? nullptr : pB + 4 * (size_t)PV_HOST_CHUNK;  // null: q9 = q8 (pv_load_soa)
        pv_arm_rows_to_planes_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(h->stage_q[slot], m, finger_left, finger_right,
                                                                                  (float4*)pA, (float4*)pB, p9);
        h->launches++;
        rc = pv_launch_state_bits(h, pA, pB, p9, nullptr, m, h->stage_bits[slot], st, false, true);
        if (rc) return rc;
        PV_CUDA(h, cudaMemcpyAsync(h_bits + done / 32, h->stage_bits[slot], (size_t)((m + 31) / 32) * sizeof(uint32_t),
                                   cudaMemcpyDeviceToHost, st));
        done += m;
        slot = (slot + 1) % PV_N_STREAMS;
    }
    for (int i = 0; i < PV_N_STREAMS; ++i) PV_CUDA(h, cudaStreamSynchronize(h->streams[i]));
    return PV_OK;
}

int pv_check_edges_host(PvHandle* h, const float* h_qa, const float* h_qb, int64_t n_edges, int n_steps,
                        float resolution, uint32_t* h_bits) {
    PV_PRECHECK(h, n_edges);
    if (!h_qa || !h_qb || !h_bits) return PV_ERR_BAD_ARG;
    int rc;
    if (n_edges <= PV_HOST_SMALL) {  // host-mapped staging, one motion per warp, one verdict byte per motion
        if ((rc = pv_ensure_small(h))) return rc;
        float* ma = (float*)h->small_host;
        float* mb = ma + (size_t)PV_HOST_SMALL * 9;
        unsigned char* ok = (unsigned char*)(mb + (size_t)PV_HOST_SMALL * 9);
        memcpy(ma, h_qa, (size_t)n_edges * 9 * sizeof(float));
        memcpy(mb, h_qb, (size_t)n_edges * 9 * sizeof(float));
        if ((rc = pv_launch_edges(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ma, mb, n_edges, n_steps, resolution,
                                  nullptr, nullptr, h->streams[0], ok)))
            return rc;
        PV_CUDA(h, cudaStreamSynchronize(h->streams[0]));
        for (int64_t w = 0; w < (n_edges + 31) / 32; ++w) {
            uint32_t word = 0;
            for (int64_t k = w * 32; k < n_edges && k < (w + 1) * 32; ++k) word |= (uint32_t)(ok[k] & 1u) << (k & 31);
            h_bits[w] = word;
        }
        return PV_OK;
    }
    rc = pv_ensure_stage(h);
    if (rc) return rc;
    int64_t done = 0;
    int slot = 0;
    while (done < n_edges) {
        const int64_t m = (n_edges - done < PV_HOST_CHUNK) ? (n_edges - done) : PV_HOST_CHUNK;
        cudaStream_t st = h->streams[slot];
        PV_CUDA(h, cudaMemcpyAsync(h->stage_q[slot], h_qa + done * 9, (size_t)m * 9 * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
        PV_CUDA(h, cudaMemcpyAsync(h->stage_q2[slot], h_qb + done * 9, (size_t)m * 9 * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
        rc = pv_launch_edges(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, h->stage_q[slot],
                             h->stage_q2[slot], m, n_steps, resolution, h->stage_bits[slot], nullptr, st, nullptr);
        if (rc) return rc;
        PV_CUDA(h, cudaMemcpyAsync(h_bits + done / 32, h->stage_bits[slot], (size_t)((m + 31) / 32) * sizeof(uint32_t),
                                   cudaMemcpyDeviceToHost, st));
        done += m;
        slot = (slot + 1) % PV_N_STREAMS;
    }
    for (int i = 0; i < PV_N_STREAMS; ++i) PV_CUDA(h, cudaStreamSynchronize(h->streams[i]));
    return PV_OK;
}

int pv_sweep(PvHandle* h, uint64_t first, int64_t n, uint32_t seed, int fingers_open, uint32_t* d_bits,
             unsigned long long* d_n_valid, float* d_q_out, void* stream) {
    PV_PRECHECK(h, n);
    if (!d_bits || (first & 31ull)) {
        snprintf(h->err, sizeof(h->err), "pv_sweep: null bits buffer or `first` not a multiple of 32");
        return PV_ERR_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t words = (n + 31) / 32;
    if (h->cull == 2) {  // tile-sorted (the default)
        const int64_t chunks = (n + PV_SB_THREADS - 1) / PV_SB_THREADS;
        const int grid = (int)(chunks < (int64_t)h->sm_count ? chunks : (int64_t)h->sm_count);
#define PV_LAUNCH_SWEEP_SORTED(CARRY, OPEN, YAW_)                                                                \
    {                                                                                                            \
        typedef PvSweepSmem<(OPEN ? 7 : 9), (OPEN ? PV_SWEEP_ST7 : PV_SWEEP_ST9)> Smem_;                         \
        const unsigned bit_ = 1u << (8 + (CARRY ? 1 : 0) + (OPEN ? 2 : 0) + (YAW_ ? 4 : 0));                     \
        if (!(h->smem_attr_mask & bit_)) {                                                                       \
            PV_CUDA(h, cudaFuncSetAttribute(pv_sweep_sorted_kernel<CARRY, OPEN, YAW_>,                           \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem_)));   \
            h->smem_attr_mask |= bit_;                                                                           \
        }                                                                                                        \
        pv_sweep_sorted_kernel<CARRY, OPEN, YAW_><<<grid, PV_SB_THREADS, sizeof(Smem_), st>>>(                   \
            h->scene, first, n, seed, d_bits, d_n_valid, d_q_out, h->gather);                                    \
    }
        if (h->scene.carry) {
            if (fingers_open) PV_LAUNCH_SWEEP_SORTED(true, true, false) else PV_LAUNCH_SWEEP_SORTED(true, false, false)
        } else if (h->all_yaw) {
            if (fingers_open) PV_LAUNCH_SWEEP_SORTED(false, true, true) else PV_LAUNCH_SWEEP_SORTED(false, false, true)
        } else {
            if (fingers_open) PV_LAUNCH_SWEEP_SORTED(false, true, false) else PV_LAUNCH_SWEEP_SORTED(false, false, false)
        }
#undef PV_LAUNCH_SWEEP_SORTED
    } else {
        if (h->scene.carry) {
            int grid = pv_grid_for(h, (const void*)pv_sweep_kernel<true, true>, PV_SB_THREADS, words);
            pv_sweep_kernel<true, true><<<grid, PV_SB_THREADS, 0, st>>>(h->scene, first, n, seed, fingers_open, d_bits,
                                                                        d_n_valid, d_q_out, h->gather);
        } else {
            int grid = pv_grid_for(h, (const void*)pv_sweep_kernel<true, false>, PV_SB_THREADS, words);
            pv_sweep_kernel<true, false><<<grid, PV_SB_THREADS, 0, st>>>(h->scene, first, n, seed, fingers_open, d_bits,
                                                                         d_n_valid, d_q_out, h->gather);
        }
    }
    h->launches++;
    PV_CUDA(h, cudaGetLastError());
    return PV_OK;
}

int pv_fp32_peak(PvHandle* h, int iters, double* tflops, float* ms_out) {
    if (pv_check_handle(h)) return PV_ERR_BAD_HANDLE;
    if (iters < 1 || !tflops) return PV_ERR_BAD_ARG;
    PvDeviceGuard guard(h->device);
    float* d_out = nullptr;
    PV_CUDA(h, cudaMalloc(&d_out, sizeof(float)));
    cudaEvent_t e0, e1;
    PV_CUDA(h, cudaEventCreate(&e0));
    PV_CUDA(h, cudaEventCreate(&e1));
    const int threads = 256, blocks = h->sm_count * 8;
    cudaStream_t st = h->streams[0];
    pv_fp32_peak_kernel<<<blocks, threads, 0, st>>>(iters / 8 + 1, 1.0f, d_out);  // warm-up
    PV_CUDA(h, cudaEventRecord(e0, st));
    pv_fp32_peak_kernel<<<blocks, threads, 0, st>>>(iters, 1.0f, d_out);
    PV_CUDA(h, cudaEventRecord(e1, st));
    PV_CUDA(h, cudaEventSynchronize(e1));
    h->launches += 2;
    float ms = 0.f;
    PV_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 128.0 * (double)iters * (double)threads * (double)blocks;
    *tflops = flops / ((double)ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return PV_OK;
}

}  // extern "C"
