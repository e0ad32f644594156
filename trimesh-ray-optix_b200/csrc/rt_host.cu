// rt_host.cu — end-to-end entry point with HOST buffers (rt_host_trace_closest).
//
// The reference's public call takes CUDA tensors (triro/ray/ray_optix.py:117-146); a caller whose
// rays live in host memory pays H2D + trace + D2H back to back, the way test/performance_test.py
// moves its result to the CPU.  This entry point pipelines the three over fixed-size ray chunks on
// kSlots private streams so the PCIe copies of chunk i+1 / i-1 overlap the traversal of chunk i.
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
static int host_slots() { return (int)env_i64("TRIRO_HOST_SLOTS", kDefaultSlots, 1, kMaxSlots); }
static int64_t host_chunk() { return env_i64("TRIRO_HOST_CHUNK", kDefaultChunk, 1024, (int64_t)1 << 26); }

struct SlotLayout {
    size_t origins, directions, hit, front, tri, loc, uv, scratch, bytes;
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
    l.bytes = off;
    return l;
}
static int64_t chunk_for(int64_t nray) { const int64_t c = host_chunk(); return nray < c ? (nray > 0 ? nray : 1) : c; }
}  // namespace rt

using namespace rt;

extern "C" int rt_host_closest_sizes(int64_t nray, size_t* dev_work_bytes) {
    RT_REQUIRE(nray >= 0 && dev_work_bytes, RT_ERR_INVALID, "rt_host_closest_sizes: bad arguments");
    *dev_work_bytes = slot_layout(chunk_for(nray)).bytes * host_slots();
    return RT_OK;
}

extern "C" int rt_host_trace_closest(const void* blob, int64_t nray, const float* h_origins, int origins_broadcast,
                                     const float* h_directions, uint8_t* h_hit, uint8_t* h_front, int32_t* h_tri_idx,
                                     float* h_loc, float* h_uv, void* dev_work, size_t dev_work_bytes) {
    RT_REQUIRE(nray >= 0, RT_ERR_INVALID, "rt_host_trace_closest: negative ray count");
    if (nray == 0) return RT_OK;
    RT_REQUIRE(blob && h_origins && h_directions && h_hit && h_front && h_tri_idx && h_loc && h_uv && dev_work,
               RT_ERR_INVALID, "rt_host_trace_closest: null pointer");
    RT_REQUIRE(((uintptr_t)dev_work & 255) == 0, RT_ERR_INVALID, "rt_host_trace_closest: dev_work must be 256-byte aligned");
    const int64_t chunk = chunk_for(nray);
    const SlotLayout lay = slot_layout(chunk);
    const int kSlots = host_slots();
    RT_REQUIRE(dev_work_bytes >= lay.bytes * kSlots, RT_ERR_SIZE, "rt_host_trace_closest: dev_work too small");

    // private streams are created once per host thread and device and reused
    static thread_local cudaStream_t streams[kMaxSlots];
    static thread_local int made = 0, made_dev = -1;
    DeviceInfo dev;
    RT_REQUIRE(device_info(&dev) == RT_OK && dev.sm_count > 0, RT_ERR_CUDA, "rt_host_trace_closest: no CUDA device");
    if (made_dev != dev.device) { made = 0; made_dev = dev.device; }
    int rc = RT_OK;
    for (; made < kSlots; ++made) {
        if (cudaStreamCreateWithFlags(&streams[made], cudaStreamNonBlocking) != cudaSuccess) {
            rc = set_error(RT_ERR_CUDA, "rt_host_trace_closest: cudaStreamCreate failed");
            break;
        }
    }
    uint8_t* base = reinterpret_cast<uint8_t*>(dev_work);
    // chunk sizes ramp up (64 Ki, 128 Ki, ... up to `chunk`) so that the first results start
    // crossing PCIe almost immediately; the D2H direction is the bottleneck of the whole call
    const int64_t ramp0 = env_i64("TRIRO_HOST_RAMP", 1 << 16, 1024, chunk);
    int64_t m = 0;
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
        uint8_t* w = base + (size_t)s * lay.bytes;
        int64_t want = c < 20 ? (ramp0 << c) : chunk;
        if (want > chunk || want <= 0) want = chunk;
        m = nray - first < want ? nray - first : want;
        cudaStream_t st = streams[s];
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
        if (e != cudaSuccess) { rc = set_error(RT_ERR_CUDA, "rt_host_trace_closest: H2D copy failed: %s", cudaGetErrorString(e)); break; }
        RT_TL_MARK(1);
        rt_ray_desc rd;
        rd.nray = m;
        rd.shape[0] = 1; rd.shape[1] = 1; rd.shape[2] = m; rd.shape[3] = 3;
        rd.origins = d_o; rd.directions = d_d;
        rd.o_stride[0] = 0; rd.o_stride[1] = 0; rd.o_stride[2] = origins_broadcast ? 0 : 3; rd.o_stride[3] = 1;
        rd.d_stride[0] = 0; rd.d_stride[1] = 0; rd.d_stride[2] = 3; rd.d_stride[3] = 1;
        rc = rt_trace_closest(blob, &rd, w + lay.hit, w + lay.front, reinterpret_cast<int32_t*>(w + lay.tri),
                              reinterpret_cast<float*>(w + lay.loc), reinterpret_cast<float*>(w + lay.uv),
                              w + lay.scratch, st);
        if (rc != RT_OK) break;
        RT_TL_MARK(2);
        e = cudaMemcpyAsync(h_hit + first, w + lay.hit, (size_t)m, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_front + first, w + lay.front, (size_t)m, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_tri_idx + first, w + lay.tri, (size_t)m * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_loc + 3 * first, w + lay.loc, (size_t)m * 12, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_uv + 2 * first, w + lay.uv, (size_t)m * 8, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { rc = set_error(RT_ERR_CUDA, "rt_host_trace_closest: D2H copy failed: %s", cudaGetErrorString(e)); break; }
        RT_TL_MARK(3);
    }
    for (int i = 0; i < made && i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(streams[i]);
        if (e != cudaSuccess && rc == RT_OK) rc = set_error(RT_ERR_CUDA, "rt_host_trace_closest: %s", cudaGetErrorString(e));
    }
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
