// fs_lav2_pool.cuh -- the HDRx32 perturbation + LAv2 render kernel with lane-level refill (rows a1-a3 of SURVEY.md
// section 8).
//
// What it computes: exactly what lav2_kernel<NumHdr<float>, ...> (fs_lav2.cuh) computes per pixel -- the AT shortcut
// (LAKernel.cuh:66-89), the LA stage walk (LAKernel.cuh:91-127, GPU_LAReference.h:241-303, GPU_LAInfoDeep.h:90-123) and
// the perturbation loop with rebasing (LAKernel.cuh:130-236) -- by calling the same per-pixel functions.  Only the
// assignment of pixels to lanes differs, so the results are the same bits.
//
// Why: a pixel's three phases have very different shapes.  AT is regular (every interior pixel of a tile runs the same
// number of passes: 30.6 of 32 lanes busy on View 14), the LA walk is not (22 of 32 lanes), the perturbation tail is
// worst (10 of 32: a few pixels of a tile iterate for thousands of steps while the rest of the warp waits).  With one
// 8x4 tile per warp from start to finish, half of the issued instructions of the View 14 frame were spent at those lane
// counts (profiles/r2_lav2_view14_summary.md).
//
// How: every warp owns two small pools of pixel states in shared memory -- pixels waiting for the LA walk and pixels
// waiting for perturbation steps (7 words each for 32-bit iteration counts: pixel, three words of delta + one of
// stage / exponent, orbit index, iteration count).  The warp alternates between three regimes:
//   * AT on a fresh tile from the queue (all 32 lanes), results pushed to the LA pool;
//   * an LA session: 32 states are popped, every lane walks its own pixel; lanes that finish push their pixel to the
//     perturbation pool and, once kRefillIdle lanes are idle, all idle lanes pop new states;
//   * a perturbation session, the same way on the other pool (the warp-synchronous rounds of fs_perturb_loop.cuh:
//     scaled plain-float chunks with a float+exponent fallback); finished pixels are stored, idle lanes refill.
// A session ends when its pool is empty and a quarter of the lanes are idle (the rest is spilled back), so the next
// tile's pixels top the pools up; when the tile queue is exhausted the pools are drained.  No state is shared between
// warps: the only synchronisation is __syncwarp().
#pragma once
#include "fs_lav2.cuh"

#ifndef FS_POOL_MIN_CTAS
#define FS_POOL_MIN_CTAS 4
#endif
#ifndef FS_POOL_TILE_LA
#define FS_POOL_TILE_LA 0 // 1: LA walk per tile (lav2_stages_v2), only the perturbation tail is pooled (View 14: 4.35 ms against 5.59
                          // with both pooled and 4.19 for the tile kernel); 0: both pooled
#endif
#ifndef FS_POOL_LA_FAST
#define FS_POOL_LA_FAST 1 // 1: select-free LA step (fs_la_fast.cuh) with the as-written step as its fallback; 0: as-written only
#endif

namespace fs {
namespace pool {

#ifdef FS_POOL_DEBUG
// development build only (make dbg): where the lane-trips of the sessions go
__device__ unsigned long long fs_pool_dbg[16];
#define FS_POOL_DBG(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&::fs::pool::fs_pool_dbg[i], (unsigned long long)(v)); } while (0)
#else
#define FS_POOL_DBG(i, v) do { } while (0)
#endif

constexpr int kCap = 64;        // entries per pool and warp (a pool never holds more than 31 + 32)
constexpr int kRefillIdle = 8;  // idle lanes that trigger a refill (or end a session whose pool is empty)
constexpr unsigned kFull = 0xffffffffu;

template <class IterT> struct Layout {
    static constexpr int kI = sizeof(IterT) / 4;
    static constexpr int kWords = 5 + 2 * kI;          // pix, a, b, c, d, ref[kI], iter[kI]
    static constexpr int kPoolWords = kWords * kCap;   // word-major (SoA): lanes with consecutive slots hit consecutive banks
    static constexpr int kWarpWords = 2 * kPoolWords;  // LA pool, then perturbation pool
    static constexpr size_t kCtaBytes = (size_t)kWarpWords * 8 * sizeof(uint32_t); // 256-thread CTAs
};

FS_D unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <class IterT> struct Entry {
    uint32_t pix, a, b, c, d;
    IterT ref, iter;
};

template <class IterT> FS_D void put(uint32_t *pool, int slot, const Entry<IterT> &e) {
    pool[0 * kCap + slot] = e.pix;
    pool[1 * kCap + slot] = e.a;
    pool[2 * kCap + slot] = e.b;
    pool[3 * kCap + slot] = e.c;
    pool[4 * kCap + slot] = e.d;
    if constexpr (sizeof(IterT) == 4) {
        pool[5 * kCap + slot] = (uint32_t)e.ref;
        pool[6 * kCap + slot] = (uint32_t)e.iter;
    } else {
        pool[5 * kCap + slot] = (uint32_t)e.ref;
        pool[6 * kCap + slot] = (uint32_t)((uint64_t)e.ref >> 32);
        pool[7 * kCap + slot] = (uint32_t)e.iter;
        pool[8 * kCap + slot] = (uint32_t)((uint64_t)e.iter >> 32);
    }
}
template <class IterT> FS_D Entry<IterT> get(const uint32_t *pool, int slot) {
    Entry<IterT> e;
    e.pix = pool[0 * kCap + slot];
    e.a = pool[1 * kCap + slot];
    e.b = pool[2 * kCap + slot];
    e.c = pool[3 * kCap + slot];
    e.d = pool[4 * kCap + slot];
    if constexpr (sizeof(IterT) == 4) {
        e.ref = (IterT)pool[5 * kCap + slot];
        e.iter = (IterT)pool[6 * kCap + slot];
    } else {
        e.ref = (IterT)((uint64_t)pool[5 * kCap + slot] | ((uint64_t)pool[6 * kCap + slot] << 32));
        e.iter = (IterT)((uint64_t)pool[7 * kCap + slot] | ((uint64_t)pool[8 * kCap + slot] << 32));
    }
    return e;
}

// Warp-uniform counters; every lane calls these converged.
FS_D int push_slot(bool want, int &cnt, unsigned lt) {
    const unsigned m = __ballot_sync(kFull, want);
    const int slot = cnt + __popc(m & lt);
    cnt += __popc(m);
    return slot;
}
FS_D int pop_slot(bool want, int &cnt, unsigned lt, bool &got) {
    const unsigned m = __ballot_sync(kFull, want);
    const int rank = __popc(m & lt);
    got = want && rank < cnt;
    const int slot = cnt - 1 - rank;
    cnt -= min(__popc(m), cnt);
    return slot;
}

template <class IterT> struct Ctx {
    uint32_t *la, *po;  // this warp's pools
    int la_cnt, po_cnt; // warp-uniform
    unsigned lt;
    unsigned long long steps, steps_at, steps_la;
};

template <class IterT> FS_D void store_pixel(const Lav2Args<NumHdr<float>, IterT> &A, uint32_t pix, IterT iter) {
    const size_t cell = (size_t)(pix >> 16) * A.pitch + (pix & 0xffffu);
    A.out[cell] = iter;
    if (A.sink) A.sink[cell] = iter;
}

// ---- LA session ------------------------------------------------------------------------------------------------
// Per-lane walk = lav2_stages_hdr32 (fs_lav2.cuh): between stages (`need`) a short loop picks the next stage whose
// threshold admits dc, otherwise one trip is one LA step.
template <class IterT, Lav2Mode Mode, bool Count>
FS_D void run_la(const Lav2Args<NumHdr<float>, IterT> &A, Ctx<IterT> &P, const bool drain) {
    using Num = NumHdr<float>;
    using Real = Hdr<float>;
    using Cplx = HdrC<float>;
    using LA = LaRec<Num, IterT>;
    Cplx dz = Num::c_zero(), dc = Num::c_zero();
    Real dcn = Num::zero();
    IterT iter = 0, jr = 0, LAIndex = 0, Macro = 0;
    uint32_t stage = 0, pix = 0;
    bool need = true, has = false, fin = false;
    FS_POOL_DBG(8, 1);
#ifdef FS_POOL_DEBUG
    const long long t0 = clock64();
#endif
    for (;;) {
        const unsigned idle_m = __ballot_sync(kFull, !has || fin);
        if (idle_m == kFull || __popc(idle_m) >= kRefillIdle) {
            // retire the finished pixels: on to the perturbation pool, or done
            const bool to_po = has && fin && Mode == Lav2Mode::Full && iter < A.n_iterations;
            if (has && fin && !to_po) store_pixel<IterT>(A, pix, iter);
            {
                const int slot = push_slot(to_po, P.po_cnt, P.lt);
                if (to_po) {
                    Entry<IterT> e;
                    e.pix = pix;
                    e.a = __float_as_uint(dz.re); e.b = (uint32_t)dz.e; e.c = __float_as_uint(dz.im); e.d = (uint32_t)dz.e;
                    e.ref = jr; e.iter = iter;
                    put<IterT>(P.po, slot, e);
                }
            }
            if (fin) { has = false; fin = false; }
            // refill the idle lanes
            {
                bool got;
                const int slot = pop_slot(!has, P.la_cnt, P.lt, got);
                if (got) {
                    const Entry<IterT> e = get<IterT>(P.la, slot);
                    pix = e.pix;
                    dz.re = __uint_as_float(e.a); dz.im = __uint_as_float(e.b); dz.e = (int)e.c;
                    stage = e.d & 0x7fffffffu;
                    need = (e.d >> 31) != 0;
                    jr = e.ref; iter = e.iter;
                    const Real dcX = Num::delta_x(A.dx, (int)(pix & 0xffffu), A.centerX);
                    const Real dcY = Num::delta_y(A.dy, (int)(pix >> 16), A.centerY);
                    dc = Num::c_make(dcX, dcY);
                    dcn = cheb(dc);
                    if (!need) {
                        const StageRec<IterT> sr = A.stages[stage];
                        LAIndex = sr.LAIndex; Macro = sr.MacroItCount;
                    }
                    has = true;
                }
            }
            __syncwarp();
            const unsigned act = __ballot_sync(kFull, has);
            if (act == 0u) break;
            if (P.po_cnt >= 32 || (!drain && P.la_cnt == 0 && __popc(act) <= 32 - kRefillIdle)) {
                // spill what is left; the next session picks it up
                const int slot = push_slot(has, P.la_cnt, P.lt);
                if (has) {
                    Entry<IterT> e;
                    e.pix = pix;
                    e.a = __float_as_uint(dz.re); e.b = __float_as_uint(dz.im); e.c = (uint32_t)dz.e;
                    e.d = stage | (need ? 0x80000000u : 0u);
                    e.ref = jr; e.iter = iter;
                    put<IterT>(P.la, slot, e);
                }
                __syncwarp();
                break;
            }
        }
#ifdef FS_POOL_DEBUG
        {
            const unsigned w = __ballot_sync(kFull, has && !fin);
            FS_POOL_DBG(drain ? 2 : 0, 1);
            FS_POOL_DBG(drain ? 3 : 1, __popc(w));
        }
#endif
        if (has && !fin) {
            if (need) {
                for (;;) {
                    if (stage == 0) { fin = true; break; }
                    stage--;
                    const StageRec<IterT> sr = A.stages[stage];
                    // isLAStageInvalid  GPU_LAReference.h:241-255
                    if (ge_pr(dcn, A.las[sr.LAIndex].LAThresholdC)) continue;
                    LAIndex = sr.LAIndex; Macro = sr.MacroItCount;
                    need = false; // jr (RefIteration) becomes the position j inside the stage
                    if (!(iter < A.n_iterations)) fin = true;
                    break;
                }
            }
            if (!fin) {
                // getLA  GPU_LAReference.h:271-303: the record is fetched whole with 128-bit loads
                const LA *recp = A.las + (LAIndex + jr);
                const LA rec = ldg_rec(recp);
                const IterT l = rec.StepLength;
                bool unusable = true, rebase = false;
                Cplx ndz = dz, z = dz;
                if (iter + l <= A.n_iterations) {
                    const Cplx nref = recp[1].Ref;
#if FS_POOL_LA_FAST
                    lafast::StepOut o;
                    if (lafast::step(rec.Ref.re, rec.Ref.im, rec.Ref.e, rec.ZCoeff.re, rec.ZCoeff.im, rec.ZCoeff.e, rec.CCoeff.re,
                                     rec.CCoeff.im, rec.CCoeff.e, rec.LAThreshold.m, rec.LAThreshold.e, nref.re, nref.im, nref.e,
                                     dz.re, dz.im, dz.e, dc.re, dc.im, dc.e, o)) {
                        unusable = o.unusable;
                        rebase = o.rebase;
                        ndz.re = o.dz.re; ndz.im = o.dz.im; ndz.e = o.dz.e;
                        z.re = o.z.re; z.im = o.z.im; z.e = o.z.e;
                    } else
#endif
                    {
                        Cplx s_ndz, s_z;
                        bool s_unusable, s_rebase;
                        la_step_as_written(rec.Ref, rec.ZCoeff, rec.CCoeff, rec.LAThreshold, nref, dz, dc, s_ndz, s_z, s_unusable, s_rebase);
                        unusable = s_unusable;
                        rebase = s_rebase;
                        if (!s_unusable) { ndz = s_ndz; z = s_z; }
                    }
                }
                if (unusable) {
                    jr = rec.NextStageLAIndex; // RefIteration for the next stage / the perturbation loop
                    need = true;
                } else {
                    iter += l;
                    if (Count) P.steps_la++;
                    jr++;
                    const bool rb = rebase || jr >= Macro;
                    dz = rb ? z : ndz;
                    jr = rb ? (IterT)0 : jr;
                    if (!(iter < A.n_iterations)) fin = true;
                }
            }
        }
    }
#ifdef FS_POOL_DEBUG
    FS_POOL_DBG(drain ? 12 : 11, clock64() - t0);
#endif
}

// ---- perturbation session ----------------------------------------------------------------------------------------
// Per-lane rounds = PerturbLoop<NumHdr<float>>::run (fs_perturb_loop.cuh): a lane in scaled form runs one speculative
// chunk per round, a lane the scaled form refused takes one float+exponent step and tries again.
template <class IterT, bool Count>
FS_D void run_po(const Lav2Args<NumHdr<float>, IterT> &A, Ctx<IterT> &P, const bool drain) {
    using namespace hdr32fast;
    using Num = NumHdr<float>;
    using Real = Hdr<float>;
    const uint4 *__restrict__ orb = reinterpret_cast<const uint4 *>(A.orbit);
    const scaled::FastElem *__restrict__ tab = reinterpret_cast<const scaled::FastElem *>(A.orbit_fast);
    const IterT last = A.orbit_count - 1;
    State a, b;
    a.dxm = 0.0f; a.dym = 0.0f; a.dxe = 0; a.dye = 0; a.z = make_uint4(0, 0, 0, 0);
    scaled::Lane L;
    L.wx = L.wy = L.ax = L.ay = L.sk = L.sk2 = L.ik = L.ccx = L.ccy = 0.0f; L.k = 0;
    scaled::CRed c;
    c.xb = c.yb = 0u; c.xe = c.ye = 0;
    Real dcX = Num::zero(), dcY = Num::zero();
    IterT RefIteration = 0, iter = 0;
    uint32_t pix = 0;
    bool has = false;
    scaled::Mode mode = scaled::kDone;
    FS_POOL_DBG(9, 1);
#ifdef FS_POOL_DEBUG
    const long long t0 = clock64();
#endif
    for (;;) {
        if (mode == scaled::kTry)
            mode = (tab != nullptr && scaled::enter<IterT>(tab, c, a.dxm, a.dxe, a.dym, a.dye, RefIteration, iter, A.n_iterations, L))
                       ? scaled::kFast : scaled::kSlow;
        const unsigned fast_m = __ballot_sync(kFull, mode == scaled::kFast);
        const unsigned slow_m = __ballot_sync(kFull, mode == scaled::kSlow);
        const unsigned busy = fast_m | slow_m;
        const int idle = 32 - __popc(busy);
        if (busy == 0u || (idle >= kRefillIdle && (P.po_cnt > 0 || !drain))) {
            bool got;
            const int slot = pop_slot(!has, P.po_cnt, P.lt, got);
            if (got) {
                const Entry<IterT> e = get<IterT>(P.po, slot);
                pix = e.pix;
                a.dxm = __uint_as_float(e.a); a.dxe = (int)e.b; a.dym = __uint_as_float(e.c); a.dye = (int)e.d;
                RefIteration = e.ref; iter = e.iter;
                dcX = Num::delta_x(A.dx, (int)(pix & 0xffffu), A.centerX);
                dcY = Num::delta_y(A.dy, (int)(pix >> 16), A.centerY);
                c = scaled::reduce_c(dcX, dcY);
                has = true;
                mode = scaled::kTry;
            }
            __syncwarp();
            const unsigned act = __ballot_sync(kFull, has);
            if (act == 0u) break;
            if (!drain && P.po_cnt == 0 && __popc(act) <= 32 - kRefillIdle) {
                if (has && mode == scaled::kFast) scaled::leave(L, a.dxm, a.dxe, a.dym, a.dye);
                const int s2 = push_slot(has, P.po_cnt, P.lt);
                if (has) {
                    Entry<IterT> e;
                    e.pix = pix;
                    e.a = __float_as_uint(a.dxm); e.b = (uint32_t)a.dxe; e.c = __float_as_uint(a.dym); e.d = (uint32_t)a.dye;
                    e.ref = RefIteration; e.iter = iter;
                    put<IterT>(P.po, s2, e);
                }
                __syncwarp();
                break;
            }
            if (__ballot_sync(kFull, got)) continue; // the new lanes enter the scaled form first
        }
        FS_POOL_DBG(drain ? 6 : 4, 1);
        FS_POOL_DBG(drain ? 7 : 5, __popc(busy));
        if (mode == scaled::kFast) {
            mode = scaled::fast_round<IterT, Count>(tab, last, A.n_iterations, c, L, RefIteration, iter, a.dxm, a.dxe, a.dym,
                                                    a.dye, P.steps);
        } else if (mode == scaled::kSlow) {
            // one float+exponent step for the lanes the scaled form refused
            a.z = __ldg(orb + RefIteration);
            if (!step<IterT, Count>(a, b, orb, last, A.n_iterations, dcX, dcY, RefIteration, iter, P.steps)) mode = scaled::kDone;
            else { a = b; mode = scaled::kTry; }
        }
        if (has && mode == scaled::kDone) {
            store_pixel<IterT>(A, pix, iter);
            has = false;
        }
    }
#ifdef FS_POOL_DEBUG
    FS_POOL_DBG(drain ? 14 : 13, clock64() - t0);
#endif
}

} // namespace pool

template <class IterT, Lav2Mode Mode, bool Count>
__global__ void __launch_bounds__(256, FS_POOL_MIN_CTAS) lav2_pool_kernel(const Lav2Args<NumHdr<float>, IterT> A) {
    using Num = NumHdr<float>;
    using Real = Hdr<float>;
    using Cplx = HdrC<float>;
    using Lay = pool::Layout<IterT>;
    extern __shared__ uint32_t fs_pool_words[];

    const int lane = threadIdx.x & 31;
    const int tiles_x = (A.width + 7) >> 3;
    const int tiles_y = (((A.height + 3) >> 2) - A.shard_index + A.shard_count - 1) / A.shard_count;
    const unsigned int n_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;

    pool::Ctx<IterT> P;
    P.la = fs_pool_words + (threadIdx.x >> 5) * Lay::kWarpWords;
    P.po = P.la + Lay::kPoolWords;
    P.la_cnt = 0;
    P.po_cnt = 0;
    P.lt = pool::lanemask_lt();
    P.steps = 0; P.steps_at = 0; P.steps_la = 0;

#ifdef FS_POOL_DEBUG
    const long long t_kernel = clock64();
#endif
    TileCursor cursor;
    tile_queue_begin(cursor);
    bool tiles_left = true;
    for (;;) {
        if (P.po_cnt >= 32) { pool::run_po<IterT, Count>(A, P, false); continue; }
        if (P.la_cnt >= 32) { pool::run_la<IterT, Mode, Count>(A, P, false); continue; }
        if (tiles_left) {
            unsigned int tile;
            if (!next_tile(A.queue, cursor, n_tiles, tile)) { tiles_left = false; continue; }
            int X, Y;
            tile_origin(tile, tiles_x, tiles_y, A.shard_count, A.shard_index, X, Y);
            X += lane & 7;
            Y += lane >> 3;
            const bool live = X < A.width && Y < A.height;
            IterT iter = 0;
            Cplx dz = Num::c_zero();
            const uint32_t pix = (uint32_t)X | ((uint32_t)Y << 16);
            if constexpr (Mode == Lav2Mode::PO) {
                // straight to the perturbation pool with a zero delta (LAKernel.cuh:130-236 starts from dz = 0)
                const int slot = pool::push_slot(live, P.po_cnt, P.lt);
                if (live) {
                    pool::Entry<IterT> e;
                    e.pix = pix;
                    e.a = __float_as_uint(dz.re); e.b = (uint32_t)dz.e; e.c = __float_as_uint(dz.im); e.d = (uint32_t)dz.e;
                    e.ref = 0; e.iter = 0;
                    pool::put<IterT>(P.po, slot, e);
                }
            } else {
                const Real dcX = Num::delta_x(A.dx, X, A.centerX);
                const Real dcY = Num::delta_y(A.dy, Y, A.centerY);
                const Cplx dc = Num::c_make(dcX, dcY);
                if (live) lav2_at<Num, IterT, Count>(A, dc, dz, iter, P.steps_at);
#if FS_POOL_TILE_LA
                // hybrid: the LA walk stays with the tile (its lanes share records: one L1 wavefront per load), only the
                // perturbation tail is pooled
                IterT RefIteration = 0;
                if constexpr (sizeof(IterT) == 4) {
                    if (live && iter < A.n_iterations && A.las2 != nullptr) lav2_stages_v2<Count>(A, dc, dz, RefIteration, iter, P.steps_la);
                    else if (live && iter < A.n_iterations) lav2_stages<Num, IterT, Count>(A, dc, dz, RefIteration, iter, P.steps_la);
                } else {
                    if (live && iter < A.n_iterations) lav2_stages<Num, IterT, Count>(A, dc, dz, RefIteration, iter, P.steps_la);
                }
                const bool to_po = live && Mode == Lav2Mode::Full && iter < A.n_iterations;
                if (live && !to_po) pool::store_pixel<IterT>(A, pix, iter);
                {
                    const int pslot = pool::push_slot(to_po, P.po_cnt, P.lt);
                    if (to_po) {
                        pool::Entry<IterT> e;
                        e.pix = pix;
                        e.a = __float_as_uint(dz.re); e.b = (uint32_t)dz.e; e.c = __float_as_uint(dz.im); e.d = (uint32_t)dz.e;
                        e.ref = RefIteration; e.iter = iter;
                        pool::put<IterT>(P.po, pslot, e);
                    }
                }
                const bool to_la = false;
#else
                const bool to_la = live && iter < A.n_iterations;
                if (live && !to_la) pool::store_pixel<IterT>(A, pix, iter);
#endif
                const int slot = pool::push_slot(to_la, P.la_cnt, P.lt);
                if (to_la) {
                    pool::Entry<IterT> e;
                    e.pix = pix;
                    e.a = __float_as_uint(dz.re); e.b = __float_as_uint(dz.im); e.c = (uint32_t)dz.e;
                    e.d = (uint32_t)(A.la_valid ? A.la_stage_count : 0) | 0x80000000u;
                    e.ref = 0; e.iter = iter;
                    pool::put<IterT>(P.la, slot, e);
                }
            }
            __syncwarp();
            continue;
        }
        if (P.la_cnt > 0) { pool::run_la<IterT, Mode, Count>(A, P, true); continue; }
        if (P.po_cnt > 0) { pool::run_po<IterT, Count>(A, P, true); continue; }
        break;
    }

#ifdef FS_POOL_DEBUG
    FS_POOL_DBG(15, clock64() - t_kernel);
#endif
    if (Count && A.step_counter) {
        unsigned long long steps = P.steps + P.steps_at + P.steps_la, steps_at = P.steps_at, steps_la = P.steps_la;
        for (int o = 16; o > 0; o >>= 1) {
            steps += __shfl_down_sync(0xffffffffu, steps, o);
            steps_at += __shfl_down_sync(0xffffffffu, steps_at, o);
            steps_la += __shfl_down_sync(0xffffffffu, steps_la, o);
        }
        if (lane == 0 && steps) atomicAdd(A.step_counter, steps);
        if (lane == 0 && steps_at) atomicAdd(A.step_counter + 1, steps_at);
        if (lane == 0 && steps_la) atomicAdd(A.step_counter + 2, steps_la);
    }
}

} // namespace fs
