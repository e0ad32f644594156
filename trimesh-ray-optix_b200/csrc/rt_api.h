// rt_api.h — shared plumbing of the C ABI translation units: error reporting, blob header
// access, launch geometry.  Nothing here is exported.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/raymesh_b200.h"
#include "rt_core.cuh"

namespace rt {

// thread-local message returned by rt_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define RT_CUDA_TRY(expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return ::rt::set_error(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                   __FILE__, __LINE__);                                           \
    } while (0)

#define RT_REQUIRE(cond, code, ...)                                \
    do {                                                           \
        if (!(cond)) return ::rt::set_error((code), __VA_ARGS__);  \
    } while (0)

struct DeviceInfo { int device; int sm_count; };
// cached per device; sm_count == 0 when there is no usable device
int device_info(DeviceInfo* out);

inline size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Triangles per leaf slot used by the builder: kLeafTrisDefault, or TRIRO_LEAF_TRIS=1..3 (read once per process;
// sizes, build and refit of one process always agree).
inline int leaf_tris_setting() {
    static const int v = [] {
        const char* e = getenv("TRIRO_LEAF_TRIS");
        const int x = e ? atoi(e) : kLeafTrisDefault;
        return x < 1 ? 1 : (x > kLeafMaxTris ? kLeafMaxTris : x);
    }();
    return v;
}

// Blob geometry derived from the triangle count alone (host side, no device read needed).
struct BlobLayout {
    size_t tris_offset, nodes_offset, parents_offset, total_bytes;
    uint32_t node_cap;
};
inline BlobLayout blob_layout(int64_t n_faces) {
    BlobLayout l;
    l.tris_offset = RT_BLOB_HEADER_BYTES;
    l.nodes_offset = align_up_sz(l.tris_offset + (size_t)(n_faces > 0 ? n_faces : 0) * 48u, 256);
    l.node_cap = wide_node_cap(n_faces, leaf_tris_setting());
    l.parents_offset = align_up_sz(l.nodes_offset + (size_t)l.node_cap * 80u, 256);
    l.total_bytes = l.parents_offset + align_up_sz((size_t)l.node_cap * 4u, 256);
    return l;
}

}  // namespace rt
