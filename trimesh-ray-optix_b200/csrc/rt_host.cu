// rt_host.cu — end-to-end entry points with HOST buffers (rt_host_trace_closest[_compact]).
//
// The reference's public call takes CUDA tensors (triro/ray/ray_optix.py:117-146); a caller whose
// rays live in host memory pays H2D + trace + D2H back to back, the way test/performance_test.py
// moves its result to the CPU.  These entry points pipeline the three over ray chunks on private
// streams so the PCIe copies of chunk i+1 / i-1 overlap the traversal of chunk i.  The compact
// variant also runs the scan + scatter of stream_compaction=True on the device, so only the hit
// mask and the packed rows of the rays that hit cross PCIe (the D2H direction bounds the call).
#include <stdio.h>
#include <stdlib.h>
#include "rt_api.h"

namespace rt {
constexpr int kMaxSlots = 8;
constexpr int kDefaultSlots = 3;
constexpr int64_t kDefaultChunk = 1 << 21;   // rays per chunk

static int64_t env_i64(const char* name, int64_t fallback, int64_t lo, int64_t hi) {
    const char* e = getenv(name);
    int64_t v = e ? atoll(e) : fallback;
    return v < lo ? lo : (v > hi ? hi : v);
}
// tuning knobs, read ONCE per process (experiments: tools/pcie_probe.py sets them before the first call)
static int host_slots() { static const int v = (int)env_i64("TRIRO_HOST_SLOTS", kDefaultSlots, 1, kMaxSlots); return v; }
static int64_t host_chunk() { static const int64_t v = env_i64("TRIRO_HOST_CHUNK", kDefaultChunk, 1024, (int64_t)1 << 26); return v; }
static int64_t host_ramp() { static const int64_t v = env_i64("TRIRO_HOST_RAMP", 1 << 16, 1024, (int64_t)1 << 26); return v; }

size_t scan_workspace_bytes(int64_t n);

struct SlotLayout {
    size_t origins, directions, hit, front, tri, loc, uv, scratch;
    size_t c_front, c_ray, c_tri, c_loc, c_uv, scan_ws, total;   // packed rows (compact variant)
    size_t scan_ws_bytes, bytes;
};
static SlotLayout slot_layout(int64_t chunk) {
    SlotLayout l;
    size_t off = 0;
    auto take = [&](size_t b) { const size_t o = off; off += align_up_sz(b, 256); return o; };
    l.origins = take((size_t)chunk * 12);
    l.directions = take((size_t)chunk * 12);
    l.hit = take((size_t)chunk);
    l.front = take((size_t)chunk);
    l.tri = take((size_t)chunk * 4);
    l.loc = take((size_t)chunk * 12);
    l.uv = take((size_t)chunk * 8);
    l.scratch = take(RT_TRACE_SCRATCH_BYTES);
    l.c_front = take((size_t)chunk);
    l.c_ray = take((size_t)chunk * 4);
    l.c_tri = take((size_t)chunk * 4);
    l.c_loc = take((size_t)chunk * 12);
    l.c_uv = take((size_t)chunk * 8);
    l.scan_ws_bytes = scan_workspace_bytes(chunk);
    l.scan_ws = take(l.scan_ws_bytes);
    l.total = take(8);
    l.bytes = off;
    return l;
}
static int64_t chunk_for(int64_t nray) { const int64_t c = host_chunk(); return nray < c ? (nray > 0 ? nray : 1) : c; }

// private streams / events / pinned totals, created once per host thread and device
struct HostCtx {
    cudaStream_t streams[kMaxSlots];
    cudaEvent_t counted[kMaxSlots];
    long long* totals;      // pinned, kMaxSlots entries
    int made, dev;
};
static int host_ctx(HostCtx** out, int slots) {
    static thread_local HostCtx ctx = {{}, {}, nullptr, 0, -1};
    DeviceInfo dev;
    RT_REQUIRE(device_info(&dev) == RT_OK && dev.sm_count > 0, RT_ERR_CUDA, "rt_host_trace_closest: no CUDA device");
    if (ctx.dev != dev.device) { ctx.made = 0; ctx.dev = dev.device; ctx.totals = nullptr; }
    if (!ctx.totals) RT_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&ctx.totals), kMaxSlots * sizeof(long long), cudaHostAllocDefault));
    for (; ctx.made < slots; ++ctx.made) {
        RT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx.streams[ctx.made], cudaStreamNonBlocking));
        RT_CUDA_TRY(cudaEventCreateWithFlags(&ctx.counted[ctx.made], cudaEventDisableTiming));
    }
    *out = &ctx;
    return RT_OK;
}

// One implementation for both entry points; COMPACT adds scan + scatter and the deferred packed copies.
template <bool COMPACT>
static int host_trace(const char* fn, const void* blob, int64_t nray, const float* h_origins, int origins_broadcast,
                      const float* h_directions, const rt_trace_opts* opts, uint8_t* h_hit, uint8_t* h_front,
                      int32_t* h_ray_idx, int32_t* h_tri_idx, float* h_loc, float* h_uv, int64_t* n_hit_out, void* dev_work,
                      size_t dev_work_bytes) {
    RT_REQUIRE(nray >= 0, RT_ERR_INVALID, "%s: negative ray count", fn);
    if (COMPACT && n_hit_out) *n_hit_out = 0;
    if (nray == 0) return RT_OK;
    RT_REQUIRE(blob && h_origins && h_directions && h_hit && h_front && h_tri_idx && h_loc && h_uv && dev_work,
               RT_ERR_INVALID, "%s: null pointer", fn);
    RT_REQUIRE(!COMPACT || (h_ray_idx && n_hit_out), RT_ERR_INVALID, "%s: null pointer", fn);
    RT_REQUIRE(!COMPACT || nray <= 0x7fffffffll, RT_ERR_INVALID, "%s: int32 ray indices need nray < 2^31", fn);
    RT_REQUIRE(((uintptr_t)dev_work & 255) == 0, RT_ERR_INVALID, "%s: dev_work must be 256-byte aligned", fn);
    const int64_t chunk = chunk_for(nray);
    const SlotLayout lay = slot_layout(chunk);
    const int kSlots = host_slots();
    RT_REQUIRE(dev_work_bytes >= lay.bytes * kSlots, RT_ERR_SIZE, "%s: dev_work too small", fn);
    HostCtx* ctx = nullptr;
    int rc = host_ctx(&ctx, kSlots);
    if (rc != RT_OK) return rc;
    uint8_t* base = reinterpret_cast<uint8_t*>(dev_work);
    rt_trace_opts o = {};
    if (opts) o = *opts;
    o.ray_first = 0; o.ray_count = -1;
    o.flags |= RT_OPT_SCRATCH_ZEROED;            // every slot's scratch is zeroed once below; launches restore it
    for (int s = 0; s < kSlots; ++s)
        RT_CUDA_TRY(cudaMemsetAsync(base + (size_t)s * lay.bytes + lay.scratch, 0, RT_TRACE_SCRATCH_BYTES, ctx->streams[s]));

    // chunk sizes ramp up (64 Ki, 128 Ki, ... up to `chunk`) so that the first results start
    // crossing PCIe almost immediately; the D2H direction is the bottleneck of the whole call
    int64_t ramp0 = host_ramp();
    if (ramp0 > chunk) ramp0 = chunk;
    int64_t m = 0;
    int64_t rows = 0;                                   // packed rows already placed (COMPACT)
    int64_t pending_first[kMaxSlots] = {0}, pending_m[kMaxSlots] = {0};
    bool pending[kMaxSlots] = {false};
    // second half of a chunk in the compact variant: its hit total has reached the host -> copy that many packed rows
    auto drain = [&](int s) -> int {
        if (!pending[s]) return RT_OK;
        pending[s] = false;
        cudaError_t e = cudaEventSynchronize(ctx->counted[s]);
        if (e != cudaSuccess) return set_error(RT_ERR_CUDA, "%s: %s", fn, cudaGetErrorString(e));
        const int64_t h = ctx->totals[s];
        if (h < 0 || h > pending_m[s]) return set_error(RT_ERR_CUDA, "%s: bad hit total %lld", fn, (long long)h);
        if (h > 0) {
            uint8_t* w = base + (size_t)s * lay.bytes;
            cudaStream_t st = ctx->streams[s];
            e = cudaMemcpyAsync(h_front + rows, w + lay.c_front, (size_t)h, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_ray_idx + rows, w + lay.c_ray, (size_t)h * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_tri_idx + rows, w + lay.c_tri, (size_t)h * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_loc + 3 * rows, w + lay.c_loc, (size_t)h * 12, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_uv + 2 * rows, w + lay.c_uv, (size_t)h * 8, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return set_error(RT_ERR_CUDA, "%s: D2H copy failed: %s", fn, cudaGetErrorString(e));
        }
        rows += h;
        return RT_OK;
    };
#ifdef RT_HOST_TIMELINE   // debugging aid (-DRT_HOST_TIMELINE): per-chunk event timeline printed to stderr
    static cudaEvent_t tl[64][4];
    static bool tl_made = false;
    if (!tl_made) { for (auto& row : tl) for (auto& ev : row) cudaEventCreate(&ev); tl_made = true; }
    int64_t tl_m[64]; int tl_n = 0;
#define RT_TL_MARK(k) if (c < 64) cudaEventRecord(tl[c][k], st)
#else
#define RT_TL_MARK(k) ((void)0)
#endif
    for (int64_t c = 0, first = 0; rc == RT_OK && first < nray; ++c, first += m) {
        const int s = (int)(c % kSlots);
        if (COMPACT) { rc = drain(s); if (rc != RT_OK) break; }      // chunks drain in order: s is the oldest pending slot
        uint8_t* w = base + (size_t)s * lay.bytes;
        int64_t want = c < 20 ? (ramp0 << c) : chunk;
        if (want > chunk || want <= 0) want = chunk;
        m = nray - first < want ? nray - first : want;
        cudaStream_t st = ctx->streams[s];
        float* d_o = reinterpret_cast<float*>(w + lay.origins);
        float* d_d = reinterpret_cast<float*>(w + lay.directions);
        cudaError_t e = cudaSuccess;
#ifdef RT_HOST_TIMELINE
        if (c < 64) { tl_m[c] = m; tl_n = (int)c + 1; }
#endif
        RT_TL_MARK(0);
        if (origins_broadcast) e = cudaMemcpyAsync(d_o, h_origins, 12, cudaMemcpyHostToDevice, st);
        else e = cudaMemcpyAsync(d_o, h_origins + 3 * first, (size_t)m * 12, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, h_directions + 3 * first, (size_t)m * 12, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = set_error(RT_ERR_CUDA, "%s: H2D copy failed: %s", fn, cudaGetErrorString(e)); break; }
        RT_TL_MARK(1);
        rt_ray_desc rd;
        rd.nray = m;
        rd.shape[0] = 1; rd.shape[1] = 1; rd.shape[2] = m; rd.shape[3] = 3;
        rd.origins = d_o; rd.directions = d_d;
        rd.o_stride[0] = 0; rd.o_stride[1] = 0; rd.o_stride[2] = origins_broadcast ? 0 : 3; rd.o_stride[3] = 1;
        rd.d_stride[0] = 0; rd.d_stride[1] = 0; rd.d_stride[2] = 3; rd.d_stride[3] = 1;
        uint8_t* d_hit = w + lay.hit; uint8_t* d_front = w + lay.front;
        int32_t* d_tri = reinterpret_cast<int32_t*>(w + lay.tri);
        float* d_loc = reinterpret_cast<float*>(w + lay.loc); float* d_uv = reinterpret_cast<float*>(w + lay.uv);
        rc = rt_trace_closest(blob, &rd, &o, d_hit, d_front, d_tri, d_loc, d_uv, w + lay.scratch, st);
        if (rc != RT_OK) break;
        RT_TL_MARK(2);
        e = cudaMemcpyAsync(h_hit + first, d_hit, (size_t)m, cudaMemcpyDeviceToHost, st);
        if (COMPACT) {
            int64_t* d_total = reinterpret_cast<int64_t*>(w + lay.total);
            rc = rt_compact_scan(d_hit, m, w + lay.scan_ws, lay.scan_ws_bytes, d_total, st);
            if (rc == RT_OK)
                rc = rt_compact_scatter_at(d_hit, m, w + lay.scan_ws, d_front, d_tri, d_loc, d_uv, first, 4, w + lay.c_front,
                                           w + lay.c_ray, reinterpret_cast<int32_t*>(w + lay.c_tri),
                                           reinterpret_cast<float*>(w + lay.c_loc), reinterpret_cast<float*>(w + lay.c_uv), st);
            if (rc != RT_OK) break;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->totals[s], d_total, 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->counted[s], st);
            pending[s] = true; pending_first[s] = first; pending_m[s] = m;
        } else {
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_front + first, d_front, (size_t)m, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_tri_idx + first, d_tri, (size_t)m * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_loc + 3 * first, d_loc, (size_t)m * 12, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_uv + 2 * first, d_uv, (size_t)m * 8, cudaMemcpyDeviceToHost, st);
        }
        if (e != cudaSuccess) { rc = set_error(RT_ERR_CUDA, "%s: D2H copy failed: %s", fn, cudaGetErrorString(e)); break; }
        RT_TL_MARK(3);
    }
    if (COMPACT && rc == RT_OK) {
        // remaining chunks, oldest first
        int64_t first_of[kMaxSlots]; int order[kMaxSlots]; int np = 0;
        for (int s = 0; s < kSlots; ++s) if (pending[s]) { order[np] = s; first_of[np] = pending_first[s]; ++np; }
        for (int i = 0; i < np; ++i)
            for (int j = i + 1; j < np; ++j)
                if (first_of[j] < first_of[i]) { const int64_t tf = first_of[i]; first_of[i] = first_of[j]; first_of[j] = tf; const int ts = order[i]; order[i] = order[j]; order[j] = ts; }
        for (int i = 0; i < np && rc == RT_OK; ++i) rc = drain(order[i]);
    }
    for (int i = 0; i < ctx->made && i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(ctx->streams[i]);
        if (e != cudaSuccess && rc == RT_OK) rc = set_error(RT_ERR_CUDA, "%s: %s", fn, cudaGetErrorString(e));
    }
    if (COMPACT && rc == RT_OK) *n_hit_out = rows;
#ifdef RT_HOST_TIMELINE
    if (getenv("TRIRO_HOST_TIMELINE")) {
        for (int c = 0; c < tl_n; ++c) {
            float t[4];
            for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tl[0][0], tl[c][k]);
            fprintf(stderr, "chunk %2d %8lld rays: h2d %7.3f..%7.3f  kernel ..%7.3f  d2h ..%7.3f ms\n", c, (long long)tl_m[c],
                    t[0], t[1], t[2], t[3]);
        }
    }
#endif
    return rc;
}
}  // namespace rt

using namespace rt;

extern "C" int rt_host_closest_sizes(int64_t nray, size_t* dev_work_bytes) {
    RT_REQUIRE(nray >= 0 && dev_work_bytes, RT_ERR_INVALID, "rt_host_closest_sizes: bad arguments");
    *dev_work_bytes = slot_layout(chunk_for(nray)).bytes * host_slots();
    return RT_OK;
}

extern "C" int rt_host_trace_closest(const void* blob, int64_t nray, const float* h_origins, int origins_broadcast,
                                     const float* h_directions, const rt_trace_opts* opts, uint8_t* h_hit, uint8_t* h_front,
                                     int32_t* h_tri_idx, float* h_loc, float* h_uv, void* dev_work, size_t dev_work_bytes) {
    return host_trace<false>("rt_host_trace_closest", blob, nray, h_origins, origins_broadcast, h_directions, opts, h_hit,
                             h_front, nullptr, h_tri_idx, h_loc, h_uv, nullptr, dev_work, dev_work_bytes);
}

extern "C" int rt_host_trace_closest_compact(const void* blob, int64_t nray, const float* h_origins, int origins_broadcast,
                                             const float* h_directions, const rt_trace_opts* opts, uint8_t* h_hit,
                                             uint8_t* h_front, int32_t* h_ray_idx, int32_t* h_tri_idx, float* h_loc,
                                             float* h_uv, int64_t* n_hit_out, void* dev_work, size_t dev_work_bytes) {
    return host_trace<true>("rt_host_trace_closest_compact", blob, nray, h_origins, origins_broadcast, h_directions, opts,
                            h_hit, h_front, h_ray_idx, h_tri_idx, h_loc, h_uv, n_hit_out, dev_work, dev_work_bytes);
}
