// rt_trace_slots.cuh — traversal kernel for INCOHERENT batches: rays live in per-warp SLOTS, not in lanes
// (included by rt_trace.cu).
//
// Same role as k_trace / k_trace_coop: replaces optixTrace + the closest/first/any/count programs of the reference
// (triro/backend/shaders.cu:67-194).  ncu of k_trace_coop on the heightfields (short rays: 3.5 wide nodes and 0.75
// triangles per ray, profiles/r2_ncu_hf4m_coop.txt) shows the node test at only 30 % of the warp instructions; the
// rest is bookkeeping executed by the ~9 lanes whose ray ends in a given step: copying a prepared ray into the lane,
// retiring it, and triangle batches that must be cut short because a finished lane cannot take a new ray while its
// (ray, triangle) pairs are still listed.  Here a ray's record (prepared ray, running result, ray index) stays in a
// shared-memory slot owned by the WARP:
//   * 64 slots per warp, handed around through three small ring queues of slot numbers: free -> (32 rays prepared at
//     full width) -> ready -> (a lane adopts a slot: 7 shared loads) -> in flight -> (traversal done) -> finished
//     -> (all its listed pairs tested, 32 slots retired at full width) -> free;
//   * a lane that finishes its ray just drops the slot number into the `finished` queue and adopts the next ready
//     slot - it never waits for triangle tests;
//   * listed (triangle, slot) pairs are tested by all 32 lanes when the list holds `tri_threshold` pairs; everything
//     that was finished before a flush is final after it, so results are written by 32 lanes for 32 rays at a time.
// Results are bit-identical to the other schedules (closest hit = 64-bit atomicMin on (t bits, primitive index)).
#pragma once

namespace rt {

#ifndef RT_SLOTS
#define RT_SLOTS 64
#endif
#ifndef RT_SLOT_LIST
#define RT_SLOT_LIST 64
#endif
constexpr int kSlots = RT_SLOTS;            // ray slots per warp (32 in flight + finished ones waiting for their pairs); >= 64 keeps
                                            // 32 prepared rays ahead, fewer slots leave more of the SM's 256 KB to the L1 cache
constexpr int kSlotListCap = RT_SLOT_LIST;  // (triangle, slot) pairs a warp can list
static_assert(kSlots >= 40 && kSlots <= 255 && kSlotListCap >= 64, "slot schedule geometry");
constexpr int kPrepMin = kSlots >= 64 ? 32 : 16;   // free slots needed before the warp prepares more rays (up to 32 at a time)
__device__ __forceinline__ int slot_wrap(int i) { return (int)((unsigned)i % (unsigned)kSlots); }
enum SlotWord { kSwO = 0, kSwId = 3, kSwS = 6, kSwOk = 9, kSwPacked = 12, kSwWords = 13 };   // words 0..5 later hold loc, uv, front
enum SlotQueue { kQFree = 0, kQReady = 1, kQFin = 2 };

template <int MODE, bool STATS>
__global__ void __launch_bounds__(kTraceThreads, MODE == kClosest ? RT_TRACE_MIN_BLOCKS : RT_TRACE_MIN_BLOCKS_LIGHT)
k_trace_slots(const __grid_constant__ TraceParams p) {
    static_assert(MODE == kClosest || MODE == kFirst || MODE == kAny || MODE == kCount, "slot schedule: closest / first / any / count");
    constexpr bool kKey = MODE == kClosest || MODE == kFirst;
    constexpr int kWarps = kTraceThreads / 32;
    __shared__ float s_slot[kWarps][kSwWords][kSlots];
    __shared__ unsigned long long s_best[kKey ? kWarps : 1][kKey ? kSlots : 1];      // (t bits << 32) | prim
    __shared__ uint32_t s_cnt[kKey ? 1 : kWarps][kKey ? 1 : kSlots];                 // count / any flag
    __shared__ uint32_t s_rlo[kWarps][kSlots], s_rhi[kWarps][kSlots];                // ray index of the slot
    __shared__ uint2 s_pair[kWarps][kSlotListCap];                                   // (triangle record, slot)
    __shared__ uint8_t s_q[kWarps][3][kSlots];                                       // ring queues of slot numbers
    init_mask_luts();
    const rt_blob_header* hdr = reinterpret_cast<const rt_blob_header*>(p.blob);
    const uint8_t* tris = p.blob + hdr->tris_offset;
    const uint8_t* nodes = p.blob + hdr->nodes_offset;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t nray = p.nray;
    const unsigned long long key_init = (unsigned long long)__float_as_uint(p.tmax) << 32;
    float (*slots)[kSlots] = s_slot[warp];
    uint8_t (*q)[kSlots] = s_q[warp];
    LocalStack stack;
    unsigned long long st_nodes = 0, st_tris = 0, st_rays = 0, st_hits = 0;

    // warp-uniform queue state
    int free_head = 0, free_count = kSlots, ready_head = 0, ready_count = 0, fin_head = 0, fin_count = 0, fin_safe = 0;
    int n_pend = 0;
    bool exhausted = false;
    for (int i = lane; i < kSlots; i += 32) q[kQFree][i] = (uint8_t)i;
    __syncwarp();

    // lane state
    Trav tv;
    Ray ray;                               // o, 1/d, octinv, magic
    float tmax = p.tmax;
    int slot = 0;
    bool active = false, nodes_done = true;
    uint32_t ty = 0u, tx = 0u, tmask = 0u;
    trav_init(tv);
    ray.magic = p.byte_magic;

    // ---- test every listed pair with the whole warp
    auto flush = [&]() {
        __syncwarp();
        for (int base = 0; base < n_pend; base += 32) {
            const int i = base + lane;
            bool won = false;
            unsigned long long key = 0;
            int sl = 0;
            TriHit h;
            Ray t;
            float v0x = 0, v0y = 0, v0z = 0, v1x = 0, v1y = 0, v1z = 0, v2x = 0, v2y = 0, v2z = 0;
            if (i < n_pend) {
                const uint2 pr = s_pair[warp][i];
                sl = (int)pr.y;
                t.Sx = slots[kSwS][sl]; t.Sy = slots[kSwS + 1][sl]; t.Sz = slots[kSwS + 2][sl];
                t.okx = slots[kSwOk][sl]; t.oky = slots[kSwOk + 1][sl]; t.okz = slots[kSwOk + 2][sl];
                t.kzf = __float_as_int(slots[kSwPacked][sl]) & 0xff;
                const uint8_t* tp = tris + (size_t)pr.x * 48u;
                const U4 a = ldg128(tp), b = ldg128(tp + 16), c = ldg128(tp + 32);
                v0x = as_float(a.x); v0y = as_float(a.y); v0z = as_float(a.z);
                v1x = as_float(b.x); v1y = as_float(b.y); v1z = as_float(b.z);
                v2x = as_float(c.x); v2y = as_float(c.y); v2z = as_float(c.z);
                if (STATS) ++st_tris;
                if (tri_test(t, v0x, v0y, v0z, v1x, v1y, v1z, v2x, v2y, v2z, h) && h.t > 0.0f) {
                    if constexpr (kKey) {
                        key = ((unsigned long long)__float_as_uint(h.t) << 32) | (unsigned long long)(uint32_t)a.w;
                        if (key < s_best[warp][sl]) { atomicMin(&s_best[warp][sl], key); won = true; }
                    } else if constexpr (MODE == kAny) {
                        if (h.t < p.tmax) s_cnt[warp][sl] = 1u;
                    } else {
                        if (h.t < p.tmax) atomicAdd(&s_cnt[warp][sl], 1u);
                    }
                }
            }
            if constexpr (MODE == kClosest) {
                __syncwarp();
                if (won && s_best[warp][sl] == key) {      // the ray's nearest so far: this lane computes its attributes
                    const HitAttr at = tri_attr(h, v0x, v0y, v0z, v1x, v1y, v1z, v2x, v2y, v2z);
                    slots[0][sl] = at.lx; slots[1][sl] = at.ly; slots[2][sl] = at.lz;
                    slots[3][sl] = at.uv0; slots[4][sl] = at.uv1;
                    slots[5][sl] = tri_front(t, h) ? 1.0f : 0.0f;
                }
            }
        }
        __syncwarp();
        n_pend = 0;
        fin_safe = fin_count;              // everything finished so far has no untested pair left
        if (active) {
            if constexpr (kKey) tmax = __uint_as_float((uint32_t)(s_best[warp][slot] >> 32));
            if constexpr (MODE == kAny) { if (s_cnt[warp][slot]) { nodes_done = true; ty = 0u; } }   // early exit
        }
    };

    // ---- write the results of the first k finished slots (k <= 32) and recycle the slots
    auto retire = [&](int k) {
        __syncwarp();
        if (lane < k) {
            const int sl = q[kQFin][slot_wrap(fin_head + lane)];
            const int64_t r = (int64_t)(((unsigned long long)s_rhi[warp][sl] << 32) | s_rlo[warp][sl]);
            if constexpr (MODE == kClosest || MODE == kFirst) {
                const unsigned long long best = s_best[warp][sl];
                const bool hit = best != key_init;
                if (STATS) { ++st_rays; st_hits += hit; }
                if (MODE == kFirst) {
                    if (p.tri) p.tri[r] = hit ? (int32_t)(uint32_t)best : -1;
                } else if (p.hit) {
                    // miss: reference miss program shaders.cu:128-135
                    p.hit[r] = hit ? 1 : 0;
                    p.front[r] = hit ? (slots[5][sl] != 0.0f ? 1 : 0) : 0;
                    p.tri[r] = hit ? (int32_t)(uint32_t)best : -1;
                    p.loc[3 * r] = hit ? slots[0][sl] : 0.f; p.loc[3 * r + 1] = hit ? slots[1][sl] : 0.f;
                    p.loc[3 * r + 2] = hit ? slots[2][sl] : 0.f;
                    p.uv[2 * r] = hit ? slots[3][sl] : 0.f; p.uv[2 * r + 1] = hit ? slots[4][sl] : 0.f;
                }
            } else if constexpr (MODE == kAny) {
                const bool found = s_cnt[warp][sl] != 0u;
                if (STATS) { ++st_rays; st_hits += found; }
                if (p.hit) p.hit[r] = found ? 1 : 0;
            } else {
                const uint32_t c = s_cnt[warp][sl];
                if (STATS) { ++st_rays; st_hits += c > 0u; }
                if (p.count) p.count[r] = (int32_t)c;
            }
            q[kQFree][slot_wrap(free_head + free_count + lane)] = (uint8_t)sl;
        }
        fin_head = slot_wrap(fin_head + k); fin_count -= k; fin_safe -= k; free_count += k;
        __syncwarp();
    };

    for (;;) {
        // ---- 1. lanes whose ray has no node work left drop its slot into the `finished` queue
        {
            const bool fin = active && nodes_done && ty == 0u;
            const unsigned m = __ballot_sync(0xffffffffu, fin);
            if (m != 0u) {
                if (fin) { q[kQFin][slot_wrap(fin_head + fin_count + __popc(m & lt_mask))] = (uint8_t)slot; active = false; }
                fin_count += __popc(m);
            }
            if (n_pend == 0) fin_safe = fin_count;
        }
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        const int n_idle = __popc(idle);
        // idle lanes, no prepared ray, no room to prepare more / nobody can work at all
        const bool starving = n_idle > 0 && ready_count == 0 && !exhausted && free_count < kPrepMin;
        const bool stuck = n_idle == 32 && ready_count == 0 && (exhausted || free_count < kPrepMin);
        // ---- 2. flush the pair list; retire finished slots (32 at a time, fewer only when slots are needed)
        if (n_pend > 0 && (n_pend >= p.tri_threshold || n_pend == kSlotListCap || stuck || (starving && free_count + fin_safe < kPrepMin)))
            flush();
        while (fin_safe >= 32 || (fin_safe > 0 && (starving || stuck))) retire(fin_safe < 32 ? fin_safe : 32);
        // ---- 3. prepare 32 more rays at full width when idle lanes outnumber the prepared rays
        if (ready_count < n_idle && !exhausted && free_count >= kPrepMin) {
            const int batch = free_count < 32 ? free_count : 32;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.ray_counter, (unsigned long long)batch);
            base = __shfl_sync(0xffffffffu, base, 0);
            int64_t n = nray - (int64_t)base;
            if (n <= batch) exhausted = true;
            if (n > batch) n = batch;
            if (n > 0) {
                if (lane < n) {
                    const int sl = q[kQFree][slot_wrap(free_head + lane)];
                    const int64_t r = (int64_t)base + lane;
                    float ox, oy, oz, dx, dy, dz;
                    load_ray<MODE>(p, r, ox, oy, oz, dx, dy, dz);
                    Ray t;
                    ray_setup(t, ox, oy, oz, dx, dy, dz);
                    slots[kSwO][sl] = t.ox; slots[kSwO + 1][sl] = t.oy; slots[kSwO + 2][sl] = t.oz;
                    slots[kSwId][sl] = t.idx; slots[kSwId + 1][sl] = t.idy; slots[kSwId + 2][sl] = t.idz;
                    slots[kSwS][sl] = t.Sx; slots[kSwS + 1][sl] = t.Sy; slots[kSwS + 2][sl] = t.Sz;
                    slots[kSwOk][sl] = t.okx; slots[kSwOk + 1][sl] = t.oky; slots[kSwOk + 2][sl] = t.okz;
                    slots[kSwPacked][sl] = __int_as_float(t.kzf | (int)(t.octinv << 8));
                    if constexpr (kKey) s_best[warp][sl] = key_init; else s_cnt[warp][sl] = 0u;
                    s_rlo[warp][sl] = (uint32_t)r; s_rhi[warp][sl] = (uint32_t)((unsigned long long)r >> 32);
                    q[kQReady][slot_wrap(ready_head + ready_count + lane)] = (uint8_t)sl;
                }
                free_head = slot_wrap(free_head + (int)n); free_count -= (int)n; ready_count += (int)n;
                __syncwarp();
            }
        }
        // ---- 4. idle lanes adopt ready slots
        if (n_idle > 0 && ready_count > 0 && (32 - n_idle) < p.refill_threshold) {
            const int take = n_idle < ready_count ? n_idle : ready_count;
            const int my = __popc(idle & lt_mask);
            if (!active && my < take) {
                slot = q[kQReady][slot_wrap(ready_head + my)];
                ray.ox = slots[kSwO][slot]; ray.oy = slots[kSwO + 1][slot]; ray.oz = slots[kSwO + 2][slot];
                ray.idx = slots[kSwId][slot]; ray.idy = slots[kSwId + 1][slot]; ray.idz = slots[kSwId + 2][slot];
                ray.octinv = ((uint32_t)__float_as_int(slots[kSwPacked][slot]) >> 8) & 0xffu;
                tmax = p.tmax;
                trav_init(tv);
                nodes_done = false;
                active = true;
            }
            ready_head = slot_wrap(ready_head + take); ready_count -= take;
            __syncwarp();          // the slot's o / 1/d words may be overwritten by attributes from now on
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (ready_count == 0 && exhausted && fin_count == 0 && n_pend == 0) break;
            continue;
        }

        // ---- 5. node phase: one wide node per lane; hit leaf slots yield a 24-bit triangle mask
        if (active && !nodes_done && ty == 0u) {
            if (tv.gy & 0xff000000u) {
                const uint32_t hits = tv.gy;
                const int bit = 31 - clz32(hits);
                tv.gy &= ~(1u << bit);
                if (tv.gy & 0xff000000u) { stack.push(tv.sp, tv.gx, tv.gy); ++tv.sp; }
                const uint32_t cs = (uint32_t)(bit - 24) ^ ray.octinv;
                const uint32_t rel = (uint32_t)popc32(hits & ~(0xffffffffu << cs));
                const uint8_t* np = nodes + (size_t)(tv.gx + rel) * 80u;
                const U4 n0 = ldg128(np), n1 = ldg128(np + 16), n2 = ldg128(np + 32), n3 = ldg128(np + 48),
                         n4 = ldg128(np + 64);
                if (STATS) ++st_nodes;
                const uint32_t hm = node_test(ray, n0, n1, n2, n3, n4, 0.0f, tmax);
                tv.gx = n1.x;
                tv.gy = (hm & 0xff000000u) | (n0.w >> 24);
                ty = hm & 0x00ffffffu; tx = n1.y; tmask = n1.z;
            }
            if (!(tv.gy & 0xff000000u)) {
                if (tv.sp == 0) nodes_done = true;
                else { --tv.sp; stack.pop(tv.sp, tv.gx, tv.gy); }
            }
        }

        // ---- 6. list this step's (triangle, slot) pairs: one warp scan of the per-lane counts positions every lane
        if (__any_sync(0xffffffffu, ty != 0u)) {
            const int c = __popc(ty);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int pos = n_pend + incl - c;
            while (ty != 0u && pos < kSlotListCap) {
                const int b = ffs32(ty) - 1;
                ty &= ty - 1u;
                s_pair[warp][pos] = make_uint2(tx + (uint32_t)popc32(tmask & ~(0xffffffffu << b)), (uint32_t)slot);
                ++pos;
            }
            n_pend = n_pend + total < kSlotListCap ? n_pend + total : kSlotListCap;
        }
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_nodes += __shfl_xor_sync(0xffffffffu, st_nodes, o);
            st_tris += __shfl_xor_sync(0xffffffffu, st_tris, o);
            st_rays += __shfl_xor_sync(0xffffffffu, st_rays, o);
            st_hits += __shfl_xor_sync(0xffffffffu, st_hits, o);
        }
        if (lane == 0) {
            atomicAdd(&p.counters[0], st_nodes);
            atomicAdd(&p.counters[1], st_tris);
            atomicAdd(&p.counters[2], st_rays);
            atomicAdd(&p.counters[3], st_hits);
        }
    }
    release_scratch(p);
}

}  // namespace rt
