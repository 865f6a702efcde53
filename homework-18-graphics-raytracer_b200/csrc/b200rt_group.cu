// b200rt_group.cu — the multi-GPU render entry points of the C ABI (include/b200rt.h, "device groups").
//
// The reference's render loop is host code (main.rs:1086-1173): the Whitted frame is one rayon pass over the pixels,
// the stochastic pass an epoch loop whose samples are summed per pixel.  Every pixel sample is independent given
// (scene, camera, params, seed, y, x, epoch), so a group of GPUs shards the path without any exchange in the data path
// (SURVEY.md 8e):
//   * epochs  (b200rt_group_render_distributed): rank g renders epochs [g*E/G, (g+1)*E/G) of the full frame into its own
//             PhotonAccumulator buffer {sum.rgb, weight_sum} (photon.rs:9-12); ONE ncclReduce(sum) to rank 0 over NVLink
//             ends the render, and only rank 0 copies the frame to the host;
//   * rows    (b200rt_group_render_whitted): rank g renders rows [g*H/G, (g+1)*H/G); the disjoint bands are gathered on
//             rank 0 with ncclSend / ncclRecv (bitwise the single-GPU frame).
// A group is either every GPU of one process (b200rt_group_create: ncclCommInitAll, one host thread per device while a
// render runs, because the wavefront tracer's host loop polls its device) or one rank per process
// (b200rt_group_create_rank: ncclCommInitRank with a unique id the host program distributes, e.g. over MPI or a file).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): single-GPU users of libb200rt.so do not need it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200rt.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // an already loaded libnccl (e.g. the one a host framework brought) is reused: same soname
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Reduce && api.Send &&
                 api.Recv && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

constexpr uint32_t kStripRows = 16u;   // rows per strip of the row-sharded stochastic frame ("interleaved 16-row strips")

// contiguous split [g*T/G, (g+1)*T/G)  
void shard_range(uint32_t total, int rank, int world, uint32_t* begin, uint32_t* count) {
    const uint64_t b = (uint64_t)rank * total / (uint64_t)world, e = (uint64_t)(rank + 1) * total / (uint64_t)world;
    *begin = (uint32_t)b;
    *count = (uint32_t)(e - b);
}

struct Member {   // one GPU driven by this process
    int device = -1, rank = -1;
    b200rt_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* d_buf = nullptr;  size_t d_buf_bytes = 0;     // accumulators / frame
    void* d_aux = nullptr;  size_t d_aux_bytes = 0;     // primary hit ids
    int rc = B200RT_OK;
    std::string err;
    float render_ms = 0.0f;
};

}  // namespace

struct b200rt_group {
    int n_ranks = 0;
    std::vector<Member> members;
    std::string last_error;
    float last_render_ms = 0.0f;   // device time of the last render on the slowest local member (incl. the collective)
};

namespace {

int fail(Member& m, int code, const std::string& what) {
    m.rc = code;
    m.err = what;
    return code;
}
#define GCU(m, call)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) return fail(m, B200RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define GNC(m, call)                                                                                     \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) return fail(m, B200RT_ERR_NCCL, std::string(#call) + ": " + nccl().GetErrorString(r__)); \
    } while (0)

int ensure(Member& m, void** p, size_t* have, size_t need) {
    if (*p && *have >= need) return B200RT_OK;
    if (*p) { GCU(m, cudaFree(*p)); *p = nullptr; *have = 0; }
    GCU(m, cudaMalloc(p, need));
    *have = need;
    return B200RT_OK;
}

int init_member(Member& m, int device, int rank) {
    m.device = device;
    m.rank = rank;
    int rc = b200rt_create(device, &m.ctx);
    if (rc != B200RT_OK) return fail(m, rc, "b200rt_create");
    GCU(m, cudaSetDevice(device));
    GCU(m, cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    GCU(m, cudaEventCreate(&m.ev0));
    GCU(m, cudaEventCreate(&m.ev1));
    return B200RT_OK;
}

// run f(member) for every local member: inline for one, one host thread per device otherwise (the wavefront tracer's
// host loop blocks on its device, and NCCL calls of different communicators must be issued concurrently)
template <class F>
int for_members(b200rt_group* g, F f) {
    if (g->members.size() == 1) f(g->members[0]);
    else {
        std::vector<std::thread> th;
        for (Member& m : g->members) th.emplace_back([&m, &f] { f(m); });
        for (std::thread& t : th) t.join();
    }
    g->last_render_ms = 0.0f;
    for (Member& m : g->members) {
        if (m.rc != B200RT_OK) {
            g->last_error = "rank " + std::to_string(m.rank) + ": " + m.err;
            if (m.ctx && m.rc == B200RT_ERR_CUDA && m.err.empty()) g->last_error += b200rt_last_cuda_error(m.ctx);
            return m.rc;
        }
        if (m.render_ms > g->last_render_ms) g->last_render_ms = m.render_ms;
    }
    return B200RT_OK;
}

// by_rows = false: epochs sharded, one ncclReduce(sum) to rank 0;  by_rows = true: rows sharded in interleaved strips (every
// rank renders all the epochs of its rows: a frame of ONE epoch - 10 M "photons" of C5 - has nothing else to split)
int render_distributed_member(b200rt_group* g, Member& m, const b200rt_camera* cam, const b200rt_params* params,
                              uint32_t epoch_begin, uint32_t epoch_count, float* out_accum, float* d_out_accum, bool by_rows) {
    m.rc = B200RT_OK;
    m.err.clear();
    const uint32_t W = params->width, H = params->height;
    const size_t n = (size_t)W * H * 4;
    GCU(m, cudaSetDevice(m.device));
    float* d_acc = (m.rank == 0 && d_out_accum) ? d_out_accum : nullptr;
    if (!d_acc) {
        int rc = ensure(m, &m.d_buf, &m.d_buf_bytes, n * sizeof(float));
        if (rc != B200RT_OK) return rc;
        d_acc = static_cast<float*>(m.d_buf);
    }
    GCU(m, cudaEventRecord(m.ev0, m.stream));
    // a fresh frame: the accumulators start at zero on the device (nothing travels host -> device but the arguments)
    GCU(m, cudaMemsetAsync(d_acc, 0, n * sizeof(float), m.stream));
    if (!by_rows) {
        uint32_t e0, en;
        shard_range(epoch_count, m.rank, g->n_ranks, &e0, &en);
        int rc = b200rt_render_distributed_device(m.ctx, cam, params, epoch_begin + e0, en, d_acc, m.stream);
        if (rc != B200RT_OK) return fail(m, rc, std::string("b200rt_render_distributed_device: ") + b200rt_last_cuda_error(m.ctx));
        if (g->n_ranks > 1) GNC(m, nccl().Reduce(d_acc, d_acc, n, ncclFloat, ncclSum, 0, m.comm, m.stream));
    } else {
        // rows in 16-row strips, strip s to rank s % G; the accumulators of the ranks have disjoint supports (zero
        // elsewhere), so one ncclReduce(sum) assembles the frame: bitwise what one GPU renders (x + 0 = x)
        int rc;
        if (g->n_ranks == 1) rc = b200rt_render_distributed_device(m.ctx, cam, params, epoch_begin, epoch_count, d_acc, m.stream);
        else rc = b200rt_render_distributed_strips_device(m.ctx, cam, params, epoch_begin, epoch_count, d_acc, m.stream, kStripRows,
                                                          (uint32_t)g->n_ranks, (uint32_t)m.rank);
        if (rc != B200RT_OK) return fail(m, rc, std::string("b200rt_render_distributed_strips_device: ") + b200rt_last_cuda_error(m.ctx));
        if (g->n_ranks > 1) GNC(m, nccl().Reduce(d_acc, d_acc, n, ncclFloat, ncclSum, 0, m.comm, m.stream));
    }
    GCU(m, cudaEventRecord(m.ev1, m.stream));
    if (m.rank == 0 && out_accum)
        GCU(m, cudaMemcpyAsync(out_accum, d_acc, n * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
    GCU(m, cudaStreamSynchronize(m.stream));
    GCU(m, cudaEventElapsedTime(&m.render_ms, m.ev0, m.ev1));
    return B200RT_OK;
}

int render_whitted_member(b200rt_group* g, Member& m, const b200rt_camera* cam, const b200rt_params* params, float* out_rgb,
                          int32_t* out_prim, float* d_out_rgb, int want_prim) {
    m.rc = B200RT_OK;
    m.err.clear();
    const uint32_t W = params->width, H = params->height;
    const uint32_t r_base = params->row_count ? params->row_begin : 0u, r_total = params->row_count ? params->row_count : H;
    const size_t npx = (size_t)W * H;
    GCU(m, cudaSetDevice(m.device));
    float* d_rgb = (m.rank == 0 && d_out_rgb) ? d_out_rgb : nullptr;
    if (!d_rgb) {
        int rc = ensure(m, &m.d_buf, &m.d_buf_bytes, npx * 3 * sizeof(float));
        if (rc != B200RT_OK) return rc;
        d_rgb = static_cast<float*>(m.d_buf);
    }
    int32_t* d_prim = nullptr;
    if (want_prim) {
        int rc = ensure(m, &m.d_aux, &m.d_aux_bytes, npx * sizeof(int32_t));
        if (rc != B200RT_OK) return rc;
        d_prim = static_cast<int32_t*>(m.d_aux);
    }
    uint32_t r0, rn;
    shard_range(r_total, m.rank, g->n_ranks, &r0, &rn);
    GCU(m, cudaEventRecord(m.ev0, m.stream));
    if (rn) {   // (row_count = 0 means "the whole frame" in b200rt_params: a rank without rows renders nothing)
        b200rt_params p = *params;
        p.row_begin = r_base + r0;
        p.row_count = rn;
        int rc = b200rt_render_whitted_device(m.ctx, cam, &p, d_rgb, d_prim, m.stream);
        if (rc != B200RT_OK) return fail(m, rc, std::string("b200rt_render_whitted_device: ") + b200rt_last_cuda_error(m.ctx));
    }
    if (g->n_ranks > 1) {
        // gather the disjoint row bands on rank 0
        GNC(m, nccl().GroupStart());
        if (m.rank == 0) {
            for (int r = 1; r < g->n_ranks; ++r) {
                uint32_t q0, qn;
                shard_range(r_total, r, g->n_ranks, &q0, &qn);
                if (!qn) continue;
                const size_t off = (size_t)(r_base + q0) * W;
                GNC(m, nccl().Recv(d_rgb + 3 * off, (size_t)qn * W * 3, ncclFloat, r, m.comm, m.stream));
                if (want_prim) GNC(m, nccl().Recv(d_prim + off, (size_t)qn * W, ncclInt32, r, m.comm, m.stream));
            }
        } else if (rn) {
            const size_t off = (size_t)(r_base + r0) * W;
            GNC(m, nccl().Send(d_rgb + 3 * off, (size_t)rn * W * 3, ncclFloat, 0, m.comm, m.stream));
            if (want_prim) GNC(m, nccl().Send(d_prim + off, (size_t)rn * W, ncclInt32, 0, m.comm, m.stream));
        }
        GNC(m, nccl().GroupEnd());
    }
    GCU(m, cudaEventRecord(m.ev1, m.stream));
    if (m.rank == 0) {
        const size_t off = (size_t)r_base * W, cnt = (size_t)r_total * W;
        if (out_rgb) GCU(m, cudaMemcpyAsync(out_rgb + 3 * off, d_rgb + 3 * off, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
        if (out_prim && want_prim) GCU(m, cudaMemcpyAsync(out_prim + off, d_prim + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, m.stream));
    }
    GCU(m, cudaStreamSynchronize(m.stream));
    GCU(m, cudaEventElapsedTime(&m.render_ms, m.ev0, m.ev1));
    return B200RT_OK;
}

bool valid_frame(const b200rt_params* p) {
    if (!p || p->width == 0 || p->height == 0) return false;
    if (p->row_count && (p->row_begin >= p->height || p->row_count > p->height - p->row_begin)) return false;
    return true;
}

}  // namespace

extern "C" {

int b200rt_group_unique_id(void* id_out, size_t id_bytes) {
    if (!id_out || id_bytes < B200RT_GROUP_ID_BYTES) return B200RT_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) <= B200RT_GROUP_ID_BYTES, "unique id size");
    if (!nccl().ok) return B200RT_ERR_NCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return B200RT_ERR_NCCL;
    std::memset(id_out, 0, id_bytes);
    std::memcpy(id_out, &id, sizeof id);
    return B200RT_OK;
}

int b200rt_group_destroy(b200rt_group* g) {
    if (!g) return B200RT_ERR_INVALID;
    for (Member& m : g->members) {
        if (m.device >= 0) cudaSetDevice(m.device);
        if (m.stream) cudaStreamSynchronize(m.stream);
        if (m.comm) nccl().CommDestroy(m.comm);
        if (m.d_buf) cudaFree(m.d_buf);
        if (m.d_aux) cudaFree(m.d_aux);
        if (m.ev0) cudaEventDestroy(m.ev0);
        if (m.ev1) cudaEventDestroy(m.ev1);
        if (m.stream) cudaStreamDestroy(m.stream);
        if (m.ctx) b200rt_destroy(m.ctx);
    }
    delete g;
    return B200RT_OK;
}

int b200rt_group_create(const int* device_ids, int n_devices, b200rt_group** out) {
    if (!out) return B200RT_ERR_INVALID;
    *out = nullptr;
    if (!device_ids || n_devices <= 0) return B200RT_ERR_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return B200RT_ERR_NO_DEVICE;
    for (int i = 0; i < n_devices; ++i) {
        if (device_ids[i] < 0 || device_ids[i] >= n) return B200RT_ERR_NO_DEVICE;
        for (int j = 0; j < i; ++j) if (device_ids[j] == device_ids[i]) return B200RT_ERR_INVALID;
    }
    if (n_devices > 1 && !nccl().ok) return B200RT_ERR_NCCL;
    b200rt_group* g = new b200rt_group();
    g->n_ranks = n_devices;
    g->members.resize((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) {
        int rc = init_member(g->members[(size_t)i], device_ids[i], i);
        if (rc != B200RT_OK) { b200rt_group_destroy(g); return rc; }
    }
    if (n_devices > 1) {
        std::vector<ncclComm_t> comms((size_t)n_devices);
        if (nccl().CommInitAll(comms.data(), n_devices, device_ids) != ncclSuccess) { b200rt_group_destroy(g); return B200RT_ERR_NCCL; }
        for (int i = 0; i < n_devices; ++i) g->members[(size_t)i].comm = comms[(size_t)i];
    }
    *out = g;
    return B200RT_OK;
}

int b200rt_group_create_rank(int device_id, int rank, int n_ranks, const void* unique_id, size_t id_bytes, b200rt_group** out) {
    if (!out) return B200RT_ERR_INVALID;
    *out = nullptr;
    if (n_ranks <= 0 || rank < 0 || rank >= n_ranks) return B200RT_ERR_INVALID;
    if (n_ranks > 1 && (!unique_id || id_bytes < sizeof(ncclUniqueId))) return B200RT_ERR_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return B200RT_ERR_NO_DEVICE;
    if (device_id < 0 || device_id >= n) return B200RT_ERR_NO_DEVICE;
    if (n_ranks > 1 && !nccl().ok) return B200RT_ERR_NCCL;
    b200rt_group* g = new b200rt_group();
    g->n_ranks = n_ranks;
    g->members.resize(1);
    int rc = init_member(g->members[0], device_id, rank);
    if (rc != B200RT_OK) { b200rt_group_destroy(g); return rc; }
    if (n_ranks > 1) {
        ncclUniqueId id;
        std::memcpy(&id, unique_id, sizeof id);
        if (cudaSetDevice(device_id) != cudaSuccess || nccl().CommInitRank(&g->members[0].comm, n_ranks, id, rank) != ncclSuccess) {
            b200rt_group_destroy(g);
            return B200RT_ERR_NCCL;
        }
    }
    *out = g;
    return B200RT_OK;
}

int b200rt_group_size(const b200rt_group* g, int* n_ranks, int* n_local) {
    if (!g) return B200RT_ERR_INVALID;
    if (n_ranks) *n_ranks = g->n_ranks;
    if (n_local) *n_local = (int)g->members.size();
    return B200RT_OK;
}

int b200rt_group_ctx(b200rt_group* g, int local_index, b200rt_ctx** out) {
    if (!g || !out || local_index < 0 || local_index >= (int)g->members.size()) return B200RT_ERR_INVALID;
    *out = g->members[(size_t)local_index].ctx;
    return B200RT_OK;
}

const char* b200rt_group_last_error(const b200rt_group* g) { return g ? g->last_error.c_str() : ""; }

int b200rt_group_upload_scene(b200rt_group* g, const b200rt_scene* scene) {
    if (!g || !scene) return B200RT_ERR_INVALID;
    for (Member& m : g->members) {   // the scene is replicated (KBs to a few MB)
        int rc = b200rt_upload_scene(m.ctx, scene);
        if (rc != B200RT_OK) { g->last_error = std::string("b200rt_upload_scene: ") + b200rt_last_cuda_error(m.ctx); return rc; }
    }
    return B200RT_OK;
}

int b200rt_group_render_distributed(b200rt_group* g, const b200rt_camera* cam, const b200rt_params* params,
                                    uint32_t epoch_begin, uint32_t epoch_count, float* out_accum) {
    if (!g || !cam || !valid_frame(params)) return B200RT_ERR_INVALID;
    bool have_root = false;
    for (const Member& m : g->members) have_root |= m.rank == 0;
    if (have_root && !out_accum) return B200RT_ERR_INVALID;
    return for_members(g, [&](Member& m) { render_distributed_member(g, m, cam, params, epoch_begin, epoch_count, out_accum, nullptr, false); });
}

int b200rt_group_render_distributed_rows(b200rt_group* g, const b200rt_camera* cam, const b200rt_params* params,
                                         uint32_t epoch_begin, uint32_t epoch_count, float* out_accum) {
    if (!g || !cam || !valid_frame(params)) return B200RT_ERR_INVALID;
    bool have_root = false;
    for (const Member& m : g->members) have_root |= m.rank == 0;
    if (have_root && !out_accum) return B200RT_ERR_INVALID;
    return for_members(g, [&](Member& m) { render_distributed_member(g, m, cam, params, epoch_begin, epoch_count, out_accum, nullptr, true); });
}

int b200rt_group_render_distributed_device(b200rt_group* g, const b200rt_camera* cam, const b200rt_params* params,
                                           uint32_t epoch_begin, uint32_t epoch_count, float* d_accum_root) {
    if (!g || !cam || !valid_frame(params)) return B200RT_ERR_INVALID;
    if (d_accum_root && ((uintptr_t)d_accum_root & 15u)) return B200RT_ERR_INVALID;
    return for_members(g, [&](Member& m) { render_distributed_member(g, m, cam, params, epoch_begin, epoch_count, nullptr, d_accum_root, false); });
}

int b200rt_group_render_distributed_rows_device(b200rt_group* g, const b200rt_camera* cam, const b200rt_params* params,
                                                uint32_t epoch_begin, uint32_t epoch_count, float* d_accum_root) {
    if (!g || !cam || !valid_frame(params)) return B200RT_ERR_INVALID;
    if (d_accum_root && ((uintptr_t)d_accum_root & 15u)) return B200RT_ERR_INVALID;
    return for_members(g, [&](Member& m) { render_distributed_member(g, m, cam, params, epoch_begin, epoch_count, nullptr, d_accum_root, true); });
}

int b200rt_group_render_whitted(b200rt_group* g, const b200rt_camera* cam, const b200rt_params* params, float* out_rgb,
                                int32_t* out_prim_id, int want_prim_ids) {
    if (!g || !cam || !valid_frame(params)) return B200RT_ERR_INVALID;
    bool have_root = false;
    for (const Member& m : g->members) have_root |= m.rank == 0;
    if (have_root && (!out_rgb || (want_prim_ids && !out_prim_id))) return B200RT_ERR_INVALID;
    const int want_prim = want_prim_ids != 0;   // every rank passes it alike: it decides whether hit ids travel
    return for_members(g, [&](Member& m) { render_whitted_member(g, m, cam, params, out_rgb, out_prim_id, nullptr, want_prim); });
}

int b200rt_group_render_whitted_device(b200rt_group* g, const b200rt_camera* cam, const b200rt_params* params, float* d_rgb_root) {
    if (!g || !cam || !valid_frame(params)) return B200RT_ERR_INVALID;
    return for_members(g, [&](Member& m) { render_whitted_member(g, m, cam, params, nullptr, nullptr, d_rgb_root, 0); });
}

int b200rt_group_last_render_ms(const b200rt_group* g, float* ms) {
    if (!g || !ms) return B200RT_ERR_INVALID;
    *ms = g->last_render_ms;
    return B200RT_OK;
}

}  // extern "C"
