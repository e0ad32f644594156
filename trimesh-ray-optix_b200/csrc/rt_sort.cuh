// rt_sort.cuh — hand-written onesweep LSD radix sort (64-bit key, 32-bit value), no CUB.
//
// Used by the BVH builder to order triangles along the Morton curve; this is one of the
// pieces the reference delegates to optixAccelBuild (triro/backend/ray.cpp:79-83).
//
// Structure (Adinets & Merrill, "Onesweep", 2022):
//   1. one histogram kernel counts all 8 digit places at once,
//   2. one tiny kernel turns the 8 x 256 counts into exclusive digit offsets,
//   3. one kernel per digit place ranks a 4096-key tile with warp-level match_any
//      multi-split, obtains the tile's global digit offsets with a decoupled look-back
//      over a chained status array (one 32-bit word per tile and digit, flag in the two
//      top bits), and scatters keys and values through shared memory so the global
//      stores are contiguous per digit run.
// The sort is stable, hence deterministic: equal Morton codes keep triangle order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {
namespace sort {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kPasses = 64 / kRadixBits;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kKeysPerThread = 16;
constexpr int kTile = kThreads * kKeysPerThread;          // 4096 keys
constexpr int kWarpSpan = 32 * kKeysPerThread;            // 512 keys per warp

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

struct Workspace {
    uint64_t* keys_alt;       // n
    uint32_t* vals_alt;       // n
    uint32_t* hist;           // kPasses * kRadix   (counts, then exclusive offsets)
    uint32_t* tile_counter;   // kPasses
    uint32_t* status;         // kPasses * n_tiles * kRadix
    int64_t n_tiles;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int64_t num_tiles(int64_t n) { return n <= 0 ? 0 : (n + kTile - 1) / kTile; }

inline size_t workspace_bytes(int64_t n) {
    const int64_t tiles = num_tiles(n);
    size_t b = 0;
    b += align_up((size_t)n * 8, 256);
    b += align_up((size_t)n * 4, 256);
    b += align_up((size_t)kPasses * kRadix * 4, 256);
    b += align_up((size_t)kPasses * 4, 256);
    b += align_up((size_t)kPasses * tiles * kRadix * 4, 256);
    return b;
}

inline Workspace carve(void* ws, int64_t n) {
    Workspace w;
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    w.n_tiles = num_tiles(n);
    w.keys_alt = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)n * 8, 256);
    w.vals_alt = reinterpret_cast<uint32_t*>(p); p += align_up((size_t)n * 4, 256);
    w.hist = reinterpret_cast<uint32_t*>(p); p += align_up((size_t)kPasses * kRadix * 4, 256);
    w.tile_counter = reinterpret_cast<uint32_t*>(p); p += align_up((size_t)kPasses * 4, 256);
    w.status = reinterpret_cast<uint32_t*>(p);
    return w;
}

// bytes from `hist` to the end of the workspace (the part that must be zero before a sort)
inline size_t zero_bytes(int64_t n) {
    return align_up((size_t)kPasses * kRadix * 4, 256) + align_up((size_t)kPasses * 4, 256) +
           align_up((size_t)kPasses * num_tiles(n) * kRadix * 4, 256);
}

// ------------------------------------------------------------------ 1. histogram of all digit places
__global__ void __launch_bounds__(256) k_histogram(const uint64_t* __restrict__ keys, int64_t n,
                                                   uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[kPasses * kRadix];
    for (int i = threadIdx.x; i < kPasses * kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < kPasses; ++p) atomicAdd(&sh[p * kRadix + (int)((k >> (p * kRadixBits)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPasses * kRadix; i += blockDim.x) {
        const uint32_t c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ------------------------------------------------------------------ 2. exclusive scan per digit place
__global__ void __launch_bounds__(kRadix) k_scan_hist(uint32_t* __restrict__ hist) {
    __shared__ uint32_t warp_sums[kRadix / 32];
    const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t c = hist[p * kRadix + t];
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int i = 0; i < w; ++i) base += warp_sums[i];
    hist[p * kRadix + t] = base + x - c;
}

// ------------------------------------------------------------------ 3. one digit place
struct PassSmem {
    uint64_t keys[kTile];
    uint32_t vals[kTile];
    uint32_t warp_hist[kWarps][kRadix];
    uint32_t digit_start[kRadix];      // tile-local start of each digit run
    uint32_t global_base[kRadix];      // global index of local position 0 of each digit run (wraps mod 2^32)
    uint32_t warp_scan[kWarps];
    uint32_t tile_id;
};

__global__ void __launch_bounds__(kThreads) k_onesweep_pass(const uint64_t* __restrict__ keys_in,
                                                            const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ vals_out, int64_t n, int pass,
                                                            const uint32_t* __restrict__ digit_offsets,
                                                            uint32_t* __restrict__ tile_counter,
                                                            volatile uint32_t* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * kRadixBits;

    // tiles are handed out in launch order so that every predecessor of a running tile is running
    if (tid == 0) sm.tile_id = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < kWarps * kRadix; i += kThreads) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile_id;
    const int64_t base = (int64_t)tile * kTile;
    const int n_tile = (int)((n - base) < (int64_t)kTile ? (n - base) : (int64_t)kTile);

    uint64_t key[kKeysPerThread];
    uint32_t val[kKeysPerThread];
    uint16_t rank[kKeysPerThread];
    const int64_t wbase = base + (int64_t)warp * kWarpSpan + lane;
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        const int64_t idx = wbase + j * 32;
        key[j] = idx < n ? keys_in[idx] : ~0ull;     // sentinels sort to the end of the (last) tile
        val[j] = idx < n ? vals_in[idx] : 0u;
    }
    // warp-level multi-split: rank of every key among the keys of its warp with the same digit
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs((int)peers) - 1;
        uint32_t pre = 0;
        if (lane == leader) {
            pre = sm.warp_hist[warp][d];
            sm.warp_hist[warp][d] = pre + (uint32_t)__popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[j] = (uint16_t)(pre + (uint32_t)__popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: exclusive scan over warps, tile count, look-back
    {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = sum;
            sum += c;
        }
        // chained scan over tiles
        volatile uint32_t* my = status + ((size_t)tile * kRadix + d);
        uint32_t excl = 0;
        if (tile == 0) {
            *my = kFlagPrefix | sum;
        } else {
            *my = kFlagAggregate | sum;
            int64_t t = (int64_t)tile - 1;
            for (;;) {
                const uint32_t s = status[(size_t)t * kRadix + d];
                const uint32_t flag = s & ~kValueMask;
                if (flag == 0u) continue;          // predecessor has not published yet
                excl += s & kValueMask;
                if (flag == kFlagPrefix) break;
                --t;
            }
            *my = kFlagPrefix | ((excl + sum) & kValueMask);
        }
        // block-wide exclusive scan of the tile counts over digits
        uint32_t x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sm.warp_scan[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int i = 0; i < warp; ++i) wb += sm.warp_scan[i];
        const uint32_t start = wb + x - sum;
        sm.digit_start[d] = start;
        sm.global_base[d] = digit_offsets[pass * kRadix + d] + excl - start;
    }
    __syncthreads();

    // scatter into shared memory in tile-sorted order
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        const uint32_t d = (uint32_t)(key[j] >> shift) & 0xffu;
        const uint32_t pos = sm.digit_start[d] + sm.warp_hist[warp][d] + rank[j];
        sm.keys[pos] = key[j];
        sm.vals[pos] = val[j];
    }
    __syncthreads();
    // contiguous runs out to global memory
    for (int i = tid; i < n_tile; i += kThreads) {
        const uint64_t k = sm.keys[i];
        const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
        const uint32_t dst = sm.global_base[d] + (uint32_t)i;
        keys_out[dst] = k;
        vals_out[dst] = sm.vals[i];
    }
}

// Sorts (keys, vals) in place; `ws` must hold workspace_bytes(n).  All launches go to `stream`.
// `passes` (<= kPasses): only the low 8 * passes key bits are sorted on - the caller guarantees the rest is zero.  With an odd
// number of passes the sorted data ends in the workspace's alternate buffers; `keys_out` / `vals_out` (optional) say where
// it is, and an odd count without them is refused.
inline cudaError_t sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, void* ws, int sm_count,
                              cudaStream_t stream, int passes = kPasses, uint64_t** keys_out = nullptr,
                              uint32_t** vals_out = nullptr) {
    if (keys_out) *keys_out = keys;
    if (vals_out) *vals_out = vals;
    if (n <= 1) return cudaSuccess;
    if (passes < 1 || passes > kPasses || ((passes & 1) && !(keys_out && vals_out))) return cudaErrorInvalidValue;
    Workspace w = carve(ws, n);
    cudaError_t e = cudaMemsetAsync(w.hist, 0, zero_bytes(n), stream);
    if (e != cudaSuccess) return e;
    int hist_blocks = (int)((n + 256 * 16 - 1) / (256 * 16));
    const int max_blocks = sm_count > 0 ? sm_count * 8 : 1184;
    if (hist_blocks > max_blocks) hist_blocks = max_blocks;
    if (hist_blocks < 1) hist_blocks = 1;
    k_histogram<<<hist_blocks, 256, 0, stream>>>(keys, n, w.hist);
    k_scan_hist<<<kPasses, kRadix, 0, stream>>>(w.hist);
    // opt-in to > 48 KB of dynamic shared memory: a per-device (per-context) attribute, so it is set on every call
    // (cheap) rather than once per process - a second GPU in the same process needs it too
    e = cudaFuncSetAttribute(k_onesweep_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem));
    if (e != cudaSuccess) return e;
    uint64_t* kin = keys; uint32_t* vin = vals;
    uint64_t* kout = w.keys_alt; uint32_t* vout = w.vals_alt;
    for (int p = 0; p < passes; ++p) {
        k_onesweep_pass<<<(unsigned)w.n_tiles, kThreads, sizeof(PassSmem), stream>>>(
            kin, vin, kout, vout, n, p, w.hist, w.tile_counter + p, w.status + (size_t)p * w.n_tiles * kRadix);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    // after the last swap (kin, vin) is where the sorted data lives: (keys, vals) for an even number of passes
    if (keys_out) *keys_out = kin;
    if (vals_out) *vals_out = vin;
    return cudaGetLastError();
}

}  // namespace sort
}  // namespace rt
