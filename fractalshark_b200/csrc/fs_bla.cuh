// fs_bla.cuh -- perturbation + bilinear-approximation (BLA) render kernels (row a4 of SURVEY.md section 8).
//
// Algorithm (what): FractalSharkGpuLib/BLAKernels.cuh:193-434 (float+exponent types: one perturbation step,
// then as many BLA skips as the table allows) and :17-168 (plain FP64: BLA skips first, then one step);
// table lookup per FractalSharkGpuLib/BLA.cuh:202-268 (`LookupBackwards`), skip evaluation per BLA.cuh:21-38.
//
// Execution (how, B200-first): the same persistent warp-tile queue as the LAv2 kernel (fs_lav2.cuh).  The
// reference's per-level arrays of 44/48/88-byte records are repacked on the device right after the upload
// into two 16-byte-aligned arrays per table -- `BlaHead {r2, l}` (16 B: what the descending validity walk
// reads, one LDG.128 per level probed) and `BlaCoef {A, B}` (32/64 B, read only for an accepted skip) --
// with the level offsets in the kernel's constant bank instead of a shared-memory pointer table filled
// behind a CTA barrier (BLAKernels.cuh:230-247).
#pragma once
#include "fs_lav2.cuh"

#ifndef FS_BLA_FAST
#define FS_BLA_FAST 1
#endif

namespace fs {

// reference wire record BLA<T> (BLA.h:7-14); natural alignment reproduces 44 / 88 / 48 / 24 bytes
template <class Num> struct BlaWire {
    typename Num::Real r2, Ax, Ay, Bx, By;
    int32_t l;
};
static_assert(sizeof(BlaWire<NumHdr<float>>) == 44 && sizeof(BlaWire<NumHdr<double>>) == 88, "BLA<HDRFloat>");
static_assert(sizeof(BlaWire<NumPlain<double>>) == 48 && sizeof(BlaWire<NumPlain<float>>) == 24, "BLA<plain>");

template <class Num> struct alignas(16) BlaHead {
    typename Num::Real r2;
    int32_t l;
};
template <class Num> struct alignas(16) BlaCoef {
    typename Num::Real Ax, Ay, Bx, By;
};
static_assert(sizeof(BlaHead<NumHdr<float>>) == 16 && sizeof(BlaHead<NumHdr<double>>) == 32 &&
              sizeof(BlaHead<NumPlain<double>>) == 16, "BlaHead");
static_assert(sizeof(BlaCoef<NumHdr<float>>) == 32 && sizeof(BlaCoef<NumHdr<double>>) == 64 &&
              sizeof(BlaCoef<NumPlain<double>>) == 32, "BlaCoef");

constexpr int kBlaMaxLevels = 34; // LM2 <= 30 in the reference (BLA.cuh:288-386) => at most 32 levels

template <class Num, class IterT> struct BlaArgs {
    IterT *out;
    const void *orbit;
    IterT orbit_count;
    const BlaHead<Num> *heads; // all levels concatenated
    const BlaCoef<Num> *coefs;
    unsigned long long level_off[kBlaMaxLevels]; // element offset of each level in heads/coefs
    int lm2;                                     // BLAS::m_LM2
    int width, height, pitch;
    int shard_count, shard_index;
    typename Num::Real dx, dy, centerX, centerY;
    IterT n_iterations;
    TileQueue queue;
    unsigned long long *step_counter;
    int cycle_watch; // 1: cycle detection at rebase events (RebaseWatch), 0: every period is executed
    int fast;        // HDRx32: 1 = select-free loop (bla_pixel_hdr32), 0 = reference-shaped operations (bla_pixel_hdr)
};

// device-side repack of one level: wire records -> heads + coefs
template <class Num>
__global__ void __launch_bounds__(256) bla_repack_kernel(const BlaWire<Num> *__restrict__ src, unsigned long long count,
                                                         BlaHead<Num> *__restrict__ heads, BlaCoef<Num> *__restrict__ coefs) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const BlaWire<Num> w = src[i];
        BlaHead<Num> h;
        h.r2 = w.r2; h.l = w.l;
        BlaCoef<Num> c;
        c.Ax = w.Ax; c.Ay = w.Ay; c.Bx = w.Bx; c.By = w.By;
        heads[i] = h;
        coefs[i] = c;
    }
}

// LookupBackwards  BLA.cuh:202-268.  Returns the flat record index and its skip length, or false.
// `zeros` is computed from the low 32 bits of m - 1 for both iteration widths, as `__clz(__brev(k))` does there.
// On View 14 the first level probed is accepted for 9 % of the skips and 2.5 levels are probed per skip (ncu), so
// the coefficients are fetched only once a level is accepted.
template <class Num, class IterT>
FS_D bool bla_lookup(const BlaArgs<Num, IterT> &A, IterT m, typename Num::Real z2, BlaCoef<Num> &coef, int &l) {
    const IterT k = m - 1;
    const int zeros = __clz((int)__brev((unsigned int)k));
    IterT ix = (sizeof(IterT) == 4 && zeros >= 32) ? (IterT)0 : (IterT)(k >> zeros);
    int level = A.lm2 == 0 ? 0 : (zeros < A.lm2 ? zeros : A.lm2);
    for (; level >= 2; --level) {
        const unsigned long long at = A.level_off[level] + (unsigned long long)ix;
        const BlaHead<Num> h = ldg_rec(A.heads + at);
        if (lt_pr(z2, h.r2)) {
            coef = ldg_rec(A.coefs + at);
            l = h.l;
            return true;
        }
        ix = ix << 1;
    }
    return false;
}

// BLA<T>::getValue  BLA.cuh:21-38, float+exponent operators (products rounded, sums aligned with one FMA each)
template <class M>
FS_D void bla_get_value(const BlaCoef<NumHdr<M>> &b, Hdr<M> &dx, Hdr<M> &dy, Hdr<M> cx, Hdr<M> cy) {
    const Hdr<M> nx = sub(add(sub(mul(b.Ax, dx), mul(b.Ay, dy)), mul(b.Bx, cx)), mul(b.By, cy));
    const Hdr<M> ny = add(add(add(mul(b.Ax, dy), mul(b.Ay, dx)), mul(b.Bx, cy)), mul(b.By, cx));
    dx = nx;
    dy = ny;
}
// plain FP64, contraction as nvcc emitted it for the reference (sm_100a SASS of mandel_1x_double_perturb_bla):
//   t1 = dx*Ay ; t2 = dy*Ay ; t1 = fma(dy,Ax,t1) ; t2 = fma(dx,Ax,-t2) ; t1 = fma(cy,Bx,t1) ; t2 = fma(cx,Bx,t2)
//   ny = fma(cx,By,t1) ; nx = fma(-cy,By,t2)
template <class M> FS_D void bla_get_value(const BlaCoef<NumPlain<M>> &b, M &dx, M &dy, M cx, M cy) {
    M t1 = dx * b.Ay;
    M t2 = dy * b.Ay;
    t1 = fma_(dy, b.Ax, t1);
    t2 = fma_(dx, b.Ax, -t2);
    t1 = fma_(cy, b.Bx, t1);
    t2 = fma_(cx, b.Bx, t2);
    dy = fma_(cx, b.By, t1);
    dx = fma_(-cy, b.By, t2);
}

// Cycle detection at rebase events: RebaseWatch (fs_types.cuh) on {dX, dY, the norm kept for the next table lookup}.
// View 14, 2^31 - 2 iterations against a 116,695-element orbit: an interior pixel runs 18,402 periods in the reference.
// ---- one pixel, float+exponent types: mandel_1xHDR_float_perturb_bla  BLAKernels.cuh:193-434 ----------------
template <class Num, class IterT, bool Count>
FS_D IterT bla_pixel_hdr(const BlaArgs<Num, IterT> &A, int X, int Y, unsigned long long &steps) {
    using Real = typename Num::Real;
    IterT iter = 0, Ref = 0;
    const Real cX = Num::delta_x(A.dx, X, A.centerX);
    const Real cY = Num::delta_y(A.dy, Y, A.centerY);
    Real dX = Num::zero(), dY = Num::zero(), dn = Num::zero();
    const IterT count = A.orbit_count;
    RebaseWatch<Real, IterT, 3> watch(A.cycle_watch != 0);

    while (iter < A.n_iterations) {
        if (Ref == 0 && iter != 0 && watch.armed) {
            const Real state[3] = {dX, dY, dn};
            iter += watch.at_rebase(state, iter, A.n_iterations);
        }
        Real zx, zy;
        OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
        ++Ref;
        Num::perturb(dX, dY, zx, zy, cX, cY);
        OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
        const Real tX = add(zx, dX), tY = add(zy, dY);
        const Real n2 = reduced(add(mul(tX, tX), mul(tY, tY))); // operator*, not square(): BLAKernels.cuh:339
        if (Count) steps++;
        if (lt_bailout(n2) && iter < A.n_iterations) {
            dn = reduced(add(mul(dX, dX), mul(dY, dY)));
            if (lt_pr(n2, dn) || Ref >= count - 1) {
                dX = tX; dY = tY; dn = n2; Ref = 0;
            }
            ++iter;
        } else {
            break;
        }
        for (;;) {
            BlaCoef<Num> b;
            int l;
            if (!bla_lookup<Num, IterT>(A, Ref, dn, b, l)) break;
            const bool res1 = Ref + (IterT)l >= count;
            const bool res2 = iter + (IterT)l >= A.n_iterations;
            const bool res3 = Ref + (IterT)l < count - 1;
            if (res1 || res2) break;
            iter += (IterT)l;
            Ref += (IterT)l;
            bla_get_value(b, dX, dY, cX, cY);
            if (Count) steps++;
            if (res3) {
                dn = reduced(add(mul(dX, dX), mul(dY, dY)));
                continue;
            }
            // landed on the last orbit element: rebase (BLAKernels.cuh:395-426); the norm kept for the next
            // lookup is that of the ORBIT point (square_mutable: clamped exponent), as written there
            OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
            dX = add(zx, dX);
            dY = add(zy, dY);
            dn = reduced(add(mul(zx, zx), mul(zy, zy)));
            Ref = 0;
            break;
        }
    }
    return iter;
}

// ---- one pixel, float+exponent binary32, select-free ------------------------------------------------------------------
// The same loop as bla_pixel_hdr with the float+exponent sums evaluated like the LAv2 kernel's perturbation step
// (fs_perturb_loop.cuh hdr32fast): alignment scales BOTH operands -- one multiplier is exactly 1 -- so there are no operand
// selects, Reduce() is bit surgery, and the reference's per-operation special cases are not evaluated per operation:
//   * exact zeros (add/sub reset the exponent of a zero result, Reduce is a no-op on zero: HDRFloat.h:414-456, 974-1065):
//     one product of the step's mantissas is tested;
//   * operator+/- drop the smaller operand at an exponent gap >= 120 (HDRFloat.h:974-1000), the select-free form at 127:
//     a gap in [120, 127) is read off the two multiplier fields (their sum is 128..134 exactly there);
//   and a step or skip that meets either is recomputed by the reference-shaped operations of bla_pixel_hdr.  (The
//   perturbation product custom_perturb2 uses the 127 rule itself, HDRFloat.h:725-794, and z = Z' + d' has a reduced,
//   non-zero Z' as its large operand wherever the gap is that wide, so only the sums of getValue need the gap test.)
// Profile of the reference-shaped loop on View 5 (profiles/r1_other_kernels_summary.md): ALU pipe 85 % busy with exactly
// those selects and compares, FMA pipe 20 %.
namespace blafast {
using hdr32fast::align;
using hdr32fast::reduce_nz;
using hdr32fast::reduce_pos;

// (a.m, a.e) + (b.m, b.e) like hdr32fast::align, also reporting an exponent gap in [120, 127)
FS_D float align_g(float am, int ae, float bm, int be, int &E, bool &gap) {
    const int fa = __viaddmin_s32_relu(ae - be, 127, 127);
    const int fb = __viaddmin_s32_relu(be - ae, 127, 127);
    gap = gap || (unsigned)(fa + fb - 128) < 7u;
    E = max(ae, be);
    return __fmaf_rn(bm, __int_as_float(fb << 23), am * __int_as_float(fa << 23));
}
} // namespace blafast

template <class IterT, bool Count>
FS_D IterT bla_pixel_hdr32(const BlaArgs<NumHdr<float>, IterT> &A, int X, int Y, unsigned long long &steps) {
    using namespace blafast;
    using Num = NumHdr<float>;
    using Real = Hdr<float>;
    const uint4 *__restrict__ orb = reinterpret_cast<const uint4 *>(A.orbit);
    IterT iter = 0, Ref = 0;
    const Real cX = Num::delta_x(A.dx, X, A.centerX);
    const Real cY = Num::delta_y(A.dy, Y, A.centerY);
    Real dX = Num::zero(), dY = Num::zero(), dn = Num::zero();
    const IterT count = A.orbit_count;
    RebaseWatch<Real, IterT, 3> watch(A.cycle_watch != 0);

    while (iter < A.n_iterations) {
        if (Ref == 0 && iter != 0 && watch.armed) {
            const Real state[3] = {dX, dY, dn};
            iter += watch.at_rebase(state, iter, A.n_iterations);
        }
        // ---- one perturbation step (BLAKernels.cuh:300-360), as hdr32fast::step evaluates it ----
        const uint4 z = __ldg(orb + Ref);
        ++Ref;
        const uint4 zn = __ldg(orb + Ref);
        Real nX, nY, tX, tY, n2, d2;
        {
            const float zxm = __uint_as_float(z.x), zym = __uint_as_float(z.w);
            int s2e, s1e;
            const float s2m = align(zxm, (int)z.y + 1, dX.m, dX.e, s2e); // 2Zx + dx
            const float s1m = align(zym, (int)z.z + 1, dY.m, dY.e, s1e); // 2Zy + dy
            const float pXa = dX.m * s2m, pXb = dY.m * s1m;
            int eX, eY;
            const float sumX = align(pXa, dX.e + s2e, -pXb, dY.e + s1e, eX);
            nX.m = align(sumX, eX, cX.m, cX.e, nX.e);
            const float pYa = dX.m * s1m, pYb = dY.m * s2m;
            const float sumY = align(pYa, dX.e + s1e, pYb, dY.e + s2e, eY);
            nY.m = align(sumY, eY, cY.m, cY.e, nY.e);
            const float nz_guard = nX.m * nY.m;
            reduce_nz(nX.m, nX.e);
            reduce_nz(nY.m, nY.e);
            tX.m = align(__uint_as_float(zn.x), (int)zn.y, nX.m, nX.e, tX.e);
            tY.m = align(__uint_as_float(zn.w), (int)zn.z, nY.m, nY.e, tY.e);
            const float sqx = tX.m * tX.m, sqy = tY.m * tY.m;
            n2.m = align(sqx, tX.e + tX.e, sqy, tY.e + tY.e, n2.e);
            reduce_pos(n2.m, n2.e);
            const float sdx = nX.m * nX.m, sdy = nY.m * nY.m;
            d2.m = align(sdx, nX.e + nX.e, sdy, nY.e + nY.e, d2.e);
            reduce_pos(d2.m, d2.e);
            // one test for every exact-zero special case of the reference
            if (((pXa * pXb) * (sqx * sqy)) * nz_guard == 0.0f) {
                Real zx{zxm, (int)z.y}, zy{zym, (int)z.z};
                nX = dX; nY = dY;
                Num::perturb(nX, nY, zx, zy, cX, cY);
                tX = add(Real{__uint_as_float(zn.x), (int)zn.y}, nX);
                tY = add(Real{__uint_as_float(zn.w), (int)zn.z}, nY);
                n2 = reduced(add(mul(tX, tX), mul(tY, tY)));
                d2 = reduced(add(mul(nX, nX), mul(nY, nY)));
            }
        }
        if (Count) steps++;
        if (!(lt_bailout(n2) && iter < A.n_iterations)) break;
        if (lt_pr(n2, d2) || Ref >= count - 1) {
            dX = tX; dY = tY; dn = n2; Ref = 0;
        } else {
            dX = nX; dY = nY; dn = d2;
        }
        ++iter;
        // ---- BLA skips (BLAKernels.cuh:362-430) ----
        for (;;) {
            BlaCoef<Num> b;
            int l;
            if (!bla_lookup<Num, IterT>(A, Ref, dn, b, l)) break;
            const bool res1 = Ref + (IterT)l >= count;
            const bool res2 = iter + (IterT)l >= A.n_iterations;
            const bool res3 = Ref + (IterT)l < count - 1;
            if (res1 || res2) break;
            iter += (IterT)l;
            Ref += (IterT)l;
            {
                // getValue  BLA.cuh:21-38:  nx = ((Ax dx - Ay dy) + Bx cx) - By cy ;  ny = ((Ax dy + Ay dx) + Bx cy) + By cx
                bool gap = false;
                const float p1 = b.Ax.m * dX.m, p2 = b.Ay.m * dY.m, p3 = b.Bx.m * cX.m, p4 = b.By.m * cY.m;
                const float q1 = b.Ax.m * dY.m, q2 = b.Ay.m * dX.m, q3 = b.Bx.m * cY.m, q4 = b.By.m * cX.m;
                const int eAx = b.Ax.e + dX.e, eAy = b.Ay.e + dY.e, eBx = b.Bx.e + cX.e, eBy = b.By.e + cY.e;
                const int fAx = b.Ax.e + dY.e, fAy = b.Ay.e + dX.e, fBx = b.Bx.e + cY.e, fBy = b.By.e + cX.e;
                int e1, e2, e3, g1, g2, g3;
                const float s1 = align_g(p1, eAx, -p2, eAy, e1, gap);
                const float s2 = align_g(s1, e1, p3, eBx, e2, gap);
                const float s3 = align_g(s2, e2, -p4, eBy, e3, gap);
                const float t1 = align_g(q1, fAx, q2, fAy, g1, gap);
                const float t2 = align_g(t1, g1, q3, fBx, g2, gap);
                const float t3 = align_g(t2, g2, q4, fBy, g3, gap);
                // zero anywhere (products or sums), or a gap the reference treats differently: the reference-shaped form
                const float zero_guard = (((p1 * p2) * (p3 * p4)) * ((q1 * q2) * (q3 * q4))) * (((s1 * s2) * s3) * ((t1 * t2) * t3));
                if (gap || zero_guard == 0.0f || !(zero_guard == zero_guard)) {
                    bla_get_value(b, dX, dY, cX, cY);
                } else {
                    dX.m = s3; dX.e = e3;
                    dY.m = t3; dY.e = g3;
                }
            }
            if (Count) steps++;
            if (res3) {
                const float sdx = dX.m * dX.m, sdy = dY.m * dY.m;
                if (sdx * sdy == 0.0f) {
                    dn = reduced(add(mul(dX, dX), mul(dY, dY)));
                } else {
                    dn.m = align(sdx, dX.e + dX.e, sdy, dY.e + dY.e, dn.e);
                    reduce_pos(dn.m, dn.e);
                }
                continue;
            }
            // landed on the last orbit element: rebase (BLAKernels.cuh:395-426); reference-shaped (once per period)
            Real zx, zy;
            OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
            dX = add(zx, dX);
            dY = add(zy, dY);
            dn = reduced(add(mul(zx, zx), mul(zy, zy)));
            Ref = 0;
            break;
        }
    }
    return iter;
}

// ---- one pixel, plain FP64: mandel_1x_double_perturb_bla  BLAKernels.cuh:17-168 ------------------------------
template <class Num, class IterT, bool Count>
FS_D IterT bla_pixel_plain(const BlaArgs<Num, IterT> &A, int X, int Y, unsigned long long &steps) {
    using Real = typename Num::Real;
    IterT iter = 0, Ref = 0;
    const Real cX = Num::delta_x(A.dx, X, A.centerX);
    const Real cY = Num::delta_y(A.dy, Y, A.centerY);
    Real dX = 0, dY = 0, dn = 0;
    const IterT count = A.orbit_count;
    RebaseWatch<Real, IterT, 3> watch(A.cycle_watch != 0);

    while (iter < A.n_iterations) {
        if (Ref == 0 && iter != 0 && watch.armed) {
            const Real state[3] = {dX, dY, dn};
            iter += watch.at_rebase(state, iter, A.n_iterations);
        }
        Real zx, zy;
        bool escaped = false;
        for (;;) {
            BlaCoef<Num> b;
            int l;
            if (!bla_lookup<Num, IterT>(A, Ref, dn, b, l)) break;
            if (Ref + (IterT)l >= count) break;
            if (iter + (IterT)l >= A.n_iterations) break;
            iter += (IterT)l;
            Ref += (IterT)l;
            bla_get_value(b, dX, dY, cX, cY);
            if (Count) steps++;
            OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
            const Real tY = dY + zy, tX = dX + zx;
            const Real n2 = Num::norm2(tX, tY);
            dn = Num::norm2(dX, dY);
            if (n2 > Real(256)) { escaped = true; break; }
            if (n2 < dn || Ref >= count - 1) {
                dX = tX; dY = tY; dn = n2; Ref = 0;
            }
        }
        (void)escaped; // the reference falls through to one more plain step either way (BLAKernels.cuh:121-166)
        if (iter >= A.n_iterations) break;

        OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
        Num::perturb(dX, dY, zx, zy, cX, cY);
        ++Ref;
        OrbitIO<Num>::load(A.orbit, Ref, zx, zy);
        const Real tY = dY + zy, tX = dX + zx;
        const Real n2 = Num::norm2(tX, tY);
        dn = Num::norm2(dX, dY);
        if (Count) steps++;
        if (n2 > Real(256)) break;
        if (n2 < dn || Ref >= count - 1) {
            dX = tX; dY = tY; dn = n2; Ref = 0;
        }
        ++iter;
    }
    return iter;
}

template <class Num, class IterT, bool Count>
__global__ void __launch_bounds__(256) bla_kernel(const BlaArgs<Num, IterT> A) {
    const int lane = threadIdx.x & 31;
    const int tiles_x = (A.width + 7) >> 3;
    const int tiles_y = (((A.height + 3) >> 2) - A.shard_index + A.shard_count - 1) / A.shard_count;
    const unsigned int n_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    unsigned long long steps = 0;
    TileCursor cursor;
    tile_queue_begin(cursor);
    for (;;) {
        unsigned int tile;
        if (!next_tile(A.queue, cursor, n_tiles, tile)) break;
        int X, Y;
        tile_origin(tile, tiles_x, tiles_y, A.shard_count, A.shard_index, X, Y);
        X += lane & 7;
        Y += lane >> 3;
        if (X < A.width && Y < A.height) {
            IterT iter;
            if constexpr (Num::kHdr && sizeof(typename Num::Mant) == 4 && FS_BLA_FAST) {
                iter = A.fast ? bla_pixel_hdr32<IterT, Count>(A, X, Y, steps) : bla_pixel_hdr<Num, IterT, Count>(A, X, Y, steps);
            } else if constexpr (Num::kHdr) iter = bla_pixel_hdr<Num, IterT, Count>(A, X, Y, steps);
            else iter = bla_pixel_plain<Num, IterT, Count>(A, X, Y, steps);
            A.out[(size_t)Y * A.pitch + X] = iter;
        }
        __syncwarp();
    }
    if (Count && A.step_counter) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(A.step_counter, steps);
    }
}

} // namespace fs
