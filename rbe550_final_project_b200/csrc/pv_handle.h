// pv_handle.h -- the opaque handle behind the C-ABI (shared by pv_kernels.cu and pv_rrtc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>

#include "../../include/panda_validity.h"
#include "pv_device.cuh"

#define PV_HANDLE_MAGIC 0x50564831u
#ifndef PV_N_STREAMS
#define PV_N_STREAMS 3
#endif
#ifndef PV_HOST_CHUNK
#define PV_HOST_CHUNK (1 << 18)  // configs per pipelined chunk of the host-buffer entry points
#endif

// Fused verdict gather: where each rank's verdict words also go (peer memory over NVLink / NVSwitch multicast).
struct PvGather {
    uint32_t* const* peers;  // device array of n_peers buffer base pointers (one per rank, symmetric memory), or null
    uint32_t* mc;            // multicast address of the same buffer (NVLS): one store reaches every rank, or null
    int n_peers;
    int pad_;
    long long word_off;      // this rank's first word inside every peer's buffer
    long long word_cap;      // words this rank may write there (stores beyond it are dropped)
};

// Epilogue shared by the verdict-producing kernels: one word per warp to the local buffer and, when a gather is
// configured, straight into every rank's gather buffer -- through the NVSwitch multicast address when there is one
// (a single store, replicated by the switch), else one peer store per rank issued by lanes 0..n_peers-1.  This fuses
// the verdict all-gather into the kernel that produces the verdicts: no collective launch, no SMs taken from the
// compute kernel.  Visibility on the peers is guaranteed at kernel completion + the symmetric-memory barrier.
#ifdef __CUDACC__
__device__ __forceinline__ void pv_emit_word(uint32_t* __restrict__ bits, const PvGather& G, int64_t w, unsigned word, int lane) {
    if (lane == 0 && bits) bits[w] = word;
    if (w >= G.word_cap) return;
    if (G.mc) {
        if (lane == 0) asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(G.mc + G.word_off + w), "r"(word) : "memory");
    } else if (G.peers) {
        if (lane < G.n_peers) G.peers[lane][G.word_off + w] = word;
    }
}

#endif

struct PvHandle {
    uint32_t magic;
    int device;
    int sm_count;
    int has_scene;
    int all_yaw;  // every scene box is rotated about world z only: the kernels' YAW instantiations apply (pv_device.cuh)
    int cull;
    int edge_cert2;      // second-tier motion certificates (pv_edge_cert2_kernel); on unless the environment says PV_EDGE_CERT2=0 at pv_create
    int launch_overlap;  // state-check launches carry the programmatic-stream-serialization attribute (pv_set_launch_overlap)
    unsigned smem_attr_mask;  // which sorted-kernel instantiations already have their dynamic shared memory opt-in
    long long launches;
    PvScene scene;
    PvGather gather;
    cudaStream_t streams[PV_N_STREAMS];
    float* stage_q[PV_N_STREAMS];
    float* stage_q2[PV_N_STREAMS];
    uint32_t* stage_bits[PV_N_STREAMS];
    void* rrtc_buf;   // device arena of the batched planner (grow-only; bounded by its chunking, see pv_rrtc.cu)
    size_t rrtc_bytes;
    void* rrtc_host;  // host-mapped pinned block: queries in, per-query results out (and packed paths of small calls)
    size_t rrtc_host_bytes;
    void* rrtc_rows_host;  // pinned mirror of the packed path rows of a large chunk
    size_t rrtc_rows_host_bytes;
    int rrtc_parity;       // which of the two row cursors the next launch uses
    void* plan_host;       // host-mapped pinned staging of pv_plan_path's edge batches (end points in, verdict words out)
    size_t plan_host_bytes;
    void* small_host;      // host-mapped pinned staging of small host-buffer calls (rows in, verdict words / bytes out)
    void* ik_buf;
    size_t ik_bytes;
    cudaMemPool_t pool;  // stream-ordered scratch of the motion validator's certificate pass (created on first use)
    char err[512];
};

// The entry points select the handle's device for their own duration and put the caller's current device back
// (a pv_* call used to leave the process on the handle's device).
struct PvDeviceGuard {
    int prev = -1;
    explicit PvDeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~PvDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    PvDeviceGuard(const PvDeviceGuard&) = delete;
    PvDeviceGuard& operator=(const PvDeviceGuard&) = delete;
};

// results of one chunk of the batched planner: (first query, count, length[], first row[], iterations[], checks[],
// packed path rows of the chunk)
typedef std::function<void(int, int, const int*, const int*, const int*, const long long*, const float*)> PvRrtcSink;
int pv_rrtc_run(PvHandle* h, const float* h_starts, const float* h_goals, int n, const PvRrtcParams* params,
                const PvRrtcSink& sink, const std::function<void(cudaStream_t)>* after_first_launch);

int pv_grid_for(PvHandle* h, const void* kernel, int threads, int64_t warps_needed);
int pv_launch_edges(PvHandle* h, const float* aA, const float* aB, const float* a9, const float* bA,
                    const float* bB, const float* b9, const float* a_aos, const float* b_aos, int64_t n,
                    int n_steps, float resolution, uint32_t* d_bits, float* d_margin, cudaStream_t st,
                    unsigned char* d_ok_bytes, bool allow_gather = false);
