// rt_wavefront.h — buffers and launcher of the wavefront stochastic tracer (rt_wavefront.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "rt_types.h"

namespace b200rt {

// consumer segments of the logic kernel: what the finished cast of a path was for
enum : int {
    WF_SEG_INIT = 0,   // round 0: every slot opens its first sample (no queue: path id = index)
    WF_SEG_PRIMARY,    // primary ray (main.rs:1150)
    WF_SEG_SHADE,      // shadow rays of get_shade (main.rs:435)
    WF_SEG_BOUNCE,     // reflected / scattered / escaped ray (main.rs:564, 583, 603)
    WF_SEG_SHB,        // no cast: get_shade still has to start (depth-0 primary hit)
    WF_SEG_REFR,       // a step of get_refract (main.rs:371, 380)
    WF_SEG_COUNT,
    // round 0 with the rays-in-lanes cast: no INIT pass and no queue - the cast generates every slot's first camera ray
    // from its index, and this segment (path id = index) regenerates it, consumes the hit and opens the sample
    WF_SEG_PRIMARY0 = WF_SEG_COUNT
};

struct WfCounters {          // per round parity
    uint32_t seg[8];         // queue length per segment
    uint32_t work[5];        // cast work items per ray slot (0 = path ray, 1..4 = shadow ray of light slot s - 1)
    uint32_t pad[3];
};
struct WfControl {
    WfCounters c[2];
    uint32_t retired;        // slots that have rendered all their samples
    uint32_t pad[15];
};

constexpr int WF_STATE_ROWS = 12;   // float4 rows of path state (192 B = 6 sectors)
constexpr uint32_t WF_WORK_PER_PATH = 5u;   // cast items a path can request in one round: its path ray + 4 shadow rays
// WF_REQ_HP = 1 lays the request rows of a path out in the 64-byte units DRAM moves: unit 0 {origin, meta} {direction}
// {shadow dir 3} {-}, unit 1 {hit position, prim} {shadow dir 0} {1} {2}, so that a shadow ray of light slot 0..2 finds its
// origin and direction in ONE unit instead of two.  Measured on B200 (round 2, 16-epoch 4K batch): cast 79.3 ms against
// 79.5, shading kernels 86.1 against 82.5 (the extra row written per level costs more than the cast gains): OFF.
#ifndef WF_REQ_HP
#define WF_REQ_HP 0
#endif
#if WF_REQ_HP
constexpr int WF_REQ_ROWS = 8;      // 128 B
#else
constexpr int WF_REQ_ROWS = 6;      // path ray (2) + 4 shadow directions (96 B = 3 sectors); shadow origins from the path state
#endif
#ifndef WF_LOGIC_MIN_BLOCKS
#define WF_LOGIC_MIN_BLOCKS 2
#endif

struct WfBuffers {
    WfControl* ctl;
    float4* st;              // [n][WF_STATE_ROWS]
    float4* req;             // [n][WF_REQ_ROWS]
    float4* res;             // [n][2]
    float2* sres;            // [n][4]
    uint32_t* q;             // [2][WF_SEG_COUNT][n]
    uint32_t* work;          // [2][WF_WORK_PER_PATH][n]: one list of path ids per ray slot
    float4* sums;            // [n] = [slot][pixel]: every slot's PhotonAccumulator {sum.rgb, weight_sum} (photon.rs:9-12), dense
    uint32_t n, n_pixels, epar;
    uint32_t fused_primary;  // round 0 ran without an INIT pass: a slot's first sample starts its accumulator row
};

// Per-kernel device times of a wavefront render (filled when the caller asks for them): CUDA events on the
// launching stream around every wf_cast_kernel launch.
struct WfKernelTiming {
    std::vector<cudaEvent_t> pool;
    double cast_ms = 0.0, logic_ms = 0.0, primary_ms = 0.0;   // primary_ms: the round-0 cast that generates the camera rays
    uint64_t cast_launches = 0;
};

size_t wf_workspace_bytes(uint32_t n_paths);
size_t wf_workspace_bytes_per_path();
uint32_t wf_epochs_in_flight(uint32_t width, uint32_t height, uint32_t epoch_count, size_t hbm_bytes);
cudaError_t launch_distributed_wavefront(const DScene& sc, const DCamera& cam, const DParams& p, float* d_accum,
                                         DCounters* d_cnt, void* workspace, uint32_t n_paths, uint32_t epar, int sm_count,
                                         uint32_t* h_pinned_retired, cudaEvent_t ev_poll, cudaStream_t stream,
                                         uint32_t* rounds_out, uint32_t* launches_out, WfKernelTiming* timing);

}  // namespace b200rt
