// fs_perturb_loop.cuh -- the plain perturbation loop with rebasing (LAKernel.cuh:130-236), generic over the
// numeric policy, plus the hand-scheduled HDRx32 version.
//
// Why a special HDRx32 loop: ncu on the generic loop (profiles/r1_lav2_v0_summary.md) shows the half-rate
// ALU pipe 91 % busy with the selects/compares of the float+exponent alignment while the FMA pipe idles at
// 22 %.  The fast step below computes the SAME roundings with a different instruction mix:
//   * alignment of (a.m, a.e) + (b.m, b.e) without selects: both operands are scaled, one of the two
//     multipliers is exactly 1:  m = fma(b.m, 2^min(-d,0), a.m * 2^min(d,0)),  d = a.e - b.e.
//     Each multiplier's exponent field is clamp(127 -/+ d, 0, 127) (one 3-input add + one min/relu), and
//     lands on 0.0f exactly when the reference's getMultiplierNeg returns 0 (|d| >= 127);
//   * Reduce() is one byte-permute + one 3-input add + one LOP3;
//   * |d'|^2 is evaluated only when its exponent bound does not already decide the rebase test;
//   * the reference's zero-mantissa special cases (add/sub reset the exponent, Reduce is a no-op) are not
//     evaluated per operation: one product of the step's mantissas is tested, and a step that touched an exact
//     zero is recomputed by the generic, reference-shaped code;
//   * the loop is unrolled by two with the state ping-ponging between two register sets, so a step writes its
//     results straight into the next step's inputs (no register copies at the loop edge).
// Results are bit-identical to the generic loop by construction; tests/test_gpu_parity.py checks them against
// the reference kernels at full size.
#pragma once
#include "fs_num.cuh"
#include "fs_scaled_loop.cuh"

namespace fs {

template <class Num> struct OrbitIO; // fs_lav2.cuh

// ---- generic loop (any numeric policy) --------------------------------------------------------------------
template <class Num, class IterT, bool Count> struct PerturbLoop {
    using Real = typename Num::Real;
    FS_D static void run(bool live, const void *orbit, const void * /*orbit_fast*/, IterT orbit_count, IterT n_iterations,
                         Real dcX, Real dcY, Real &dX, Real &dY, IterT &RefIteration, IterT &iter, unsigned long long &steps) {
        if (!live) return;
        Real zx, zy;
        OrbitIO<Num>::load(orbit, RefIteration, zx, zy);
        const IterT last = orbit_count - 1;
        // (Tried for the plain types and dropped: eight speculative steps per branch with the tests folded into one
        // predicate, as in the direct and scaled kernels.  On the mid-depth views these types serve, a rebase fires
        // every few steps, most chunks are discarded, and the frame got slower: f64 LAv2 0.88x, f32 0.74x of the
        // reference kernel against 1.3-1.6x for this loop.)
        for (;;) {
            Num::perturb(dX, dY, zx, zy, dcX, dcY);
            ++RefIteration;
            OrbitIO<Num>::load(orbit, RefIteration, zx, zy);
            const Real tX = add(zx, dX);
            const Real tY = add(zy, dY);
            const Real n2 = Num::norm2(tX, tY);
            if (Count) steps++;
            if (lt_bailout(n2) && iter < n_iterations) {
                const Real d2 = Num::norm2(dX, dY);
                if (lt_pr(n2, d2) || RefIteration >= last) {
                    dX = tX;
                    dY = tY;
                    RefIteration = 0;
                    OrbitIO<Num>::load(orbit, 0, zx, zy);
                }
                ++iter;
            } else {
                break;
            }
        }
    }
};

// ---- HDRx32: select-free step ------------------------------------------------------------------------------
namespace hdr32fast {

// (a.m, a.e) + (b.m, b.e): mantissa returned, exponent max(a.e, b.e) in E
FS_D float align(float am, int ae, float bm, int be, int &E) {
    const int fa = __viaddmin_s32_relu(ae - be, 127, 127); // clamp(127 + d, 0, 127): field of 2^min(d,0)
    const int fb = __viaddmin_s32_relu(be - ae, 127, 127); // clamp(127 - d, 0, 127): field of 2^min(-d,0)
    const float ma = __int_as_float(fa << 23);
    const float mb = __int_as_float(fb << 23);
    E = max(ae, be);
    return __fmaf_rn(bm, mb, am * ma);
}
// Reduce() of a non-zero mantissa (HDRFloat.h:432-448)
FS_D void reduce_nz(float &m, int &e) {
    const unsigned b = __float_as_uint(m);
    const unsigned fe = __byte_perm(b + b, 0u, 0x4443); // exponent field: byte 3 of (bits << 1)
    e = e + (int)fe - 127;
    m = __uint_as_float((b & 0x807fffffu) | 0x3f800000u);
}
FS_D void reduce_pos(float &m, int &e) { // mantissa known positive
    const unsigned b = __float_as_uint(m);
    e = e + (int)(b >> 23) - 127;
    m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
}

struct State {
    float dxm, dym;
    int dxe, dye;
    uint4 z; // orbit element at RefIteration: {x.m, x.e, y.e, y.m}
};

// One perturbation step: reads `in`, writes `out` (the state the next step starts from).
// Returns false when the pixel is finished (escaped or out of iterations).
template <class IterT, bool Count>
FS_D bool step(const State &in, State &out, const uint4 *__restrict__ orb, IterT last, IterT n_iterations,
               Hdr<float> dcX, Hdr<float> dcY, IterT &RefIteration, IterT &iter, unsigned long long &steps) {
    using Num = NumHdr<float>;
    using Real = Hdr<float>;
    ++RefIteration;
    const uint4 zn = __ldg(orb + RefIteration);
    const float dxm = in.dxm, dym = in.dym;
    const int dxe = in.dxe, dye = in.dye;
    const float zxm = __uint_as_float(in.z.x), zym = __uint_as_float(in.z.w);
    const int zxe1 = (int)in.z.y + 1, zye1 = (int)in.z.z + 1; // 2*Z: exponent + 1 (the MIN_BIG clamp cannot trigger)
    // tempSum2 = 2Zx + dx ; tempSum1 = 2Zy + dy
    int s2e, s1e;
    const float s2m = align(zxm, zxe1, dxm, dxe, s2e);
    const float s1m = align(zym, zye1, dym, dye, s1e);
    // custom_perturb2, X: dx*s2 - dy*s1 + cx
    const float pXa = dxm * s2m, pXb = dym * s1m;
    int eX, nxe;
    const float sumX = align(pXa, dxe + s2e, -pXb, dye + s1e, eX);
    float nxm = align(sumX, eX, dcX.m, dcX.e, nxe);
    // Y: dx*s1 + dy*s2 + cy
    const float pYa = dxm * s1m, pYb = dym * s2m;
    int eY, nye;
    const float sumY = align(pYa, dxe + s1e, pYb, dye + s2e, eY);
    float nym = align(sumY, eY, dcY.m, dcY.e, nye);
    const float nz_guard = nxm * nym; // raw (unreduced) results: a zero here needs the reference's special cases
    reduce_nz(nxm, nxe);
    reduce_nz(nym, nye);
    // z = Z' + d'
    const float wxm = __uint_as_float(zn.x), wym = __uint_as_float(zn.w);
    int txe, tye;
    float txm = align(wxm, (int)zn.y, nxm, nxe, txe);
    float tym = align(wym, (int)zn.z, nym, nye, tye);
    // |z|^2
    const float sqx = txm * txm, sqy = tym * tym;
    int n2e;
    float n2m = align(sqx, txe + txe, sqy, tye + tye, n2e);
    reduce_pos(n2m, n2e);
    bool below = n2e <= 1; // reduced, non-zero: "< 256" can never fail at exponent 1 (HDRFloat.h:1169-1184)
    // |d'|^2 only matters for the rebase test |z|^2 < |d'|^2.  d'x, d'y are reduced (mantissas in [1,2)), so
    // |d'|^2 < 2^(2*max(e)+3) and its reduced exponent is <= 2*max(e)+2: when |z|^2 has a larger exponent the
    // comparison is decided without evaluating |d'|^2 (the usual case, |z| >> |d'|).
    bool rebase = false;
    if (n2e <= 2 * max(nxe, nye) + 2) {
        const float sdx = nxm * nxm, sdy = nym * nym;
        int d2e;
        float d2m = align(sdx, nxe + nxe, sdy, nye + nye, d2e);
        reduce_pos(d2m, d2e);
        rebase = n2e < d2e || (n2e == d2e && n2m < d2m);
    }
    // one test for every exact-zero special case of the reference (see file header)
    const float guard = ((pXa * pXb) * (sqx * sqy)) * nz_guard;
    if (guard == 0.0f) {
        Real dX{dxm, dxe}, dY{dym, dye};
        const Real zx{zxm, (int)in.z.y}, zy{zym, (int)in.z.z};
        Num::perturb(dX, dY, zx, zy, dcX, dcY);
        const Real tX = add(Real{wxm, (int)zn.y}, dX), tY = add(Real{wym, (int)zn.z}, dY);
        const Real n2 = Num::norm2(tX, tY), d2 = Num::norm2(dX, dY);
        nxm = dX.m; nxe = dX.e; nym = dY.m; nye = dY.e;
        txm = tX.m; txe = tX.e; tym = tY.m; tye = tY.e;
        below = lt_bailout(n2);
        rebase = lt_pr(n2, d2);
    }
    if (Count) steps++;
    if (!(below && iter < n_iterations)) return false;
    ++iter;
    if (rebase || RefIteration >= last) {
        out.dxm = txm; out.dxe = txe; out.dym = tym; out.dye = tye;
        RefIteration = 0;
        out.z = __ldg(orb);
    } else {
        out.dxm = nxm; out.dxe = nxe; out.dym = nym; out.dye = nye;
        out.z = zn;
    }
    return true;
}

} // namespace hdr32fast

template <class IterT, bool Count> struct PerturbLoop<NumHdr<float>, IterT, Count> {
    using Real = Hdr<float>;
    // orbit_fast: the per-element table of fs_scaled_loop.cuh (built on upload); nullptr selects the pure
    // float+exponent loop.
#ifndef FS_SLOW_BATCH
#define FS_SLOW_BATCH 1
#endif
    static constexpr int kSlowBatch = FS_SLOW_BATCH; // lanes that must be waiting before a float+exponent step is issued
    // Warp-synchronous: all 32 lanes call this converged; `live` = the lane has a pixel to iterate.
    FS_D static void run(bool live, const void *orbit, const void *orbit_fast, IterT orbit_count, IterT n_iterations, Real dcX,
                         Real dcY, Real &dXio, Real &dYio, IterT &RefIteration, IterT &iter, unsigned long long &steps) {
        using namespace hdr32fast;
        const uint4 *__restrict__ orb = reinterpret_cast<const uint4 *>(orbit);
        const IterT last = orbit_count - 1;
        State a, b;
        a.dxm = dXio.m; a.dxe = dXio.e; a.dym = dYio.m; a.dye = dYio.e;
        if (orbit_fast == nullptr) {
            if (!live) return;
            a.z = __ldg(orb + RefIteration);
            for (;;) {
                if (!step<IterT, Count>(a, b, orb, last, n_iterations, dcX, dcY, RefIteration, iter, steps)) break;
                if (!step<IterT, Count>(b, a, orb, last, n_iterations, dcX, dcY, RefIteration, iter, steps)) break;
            }
            return;
        }
        const scaled::FastElem *__restrict__ tab = reinterpret_cast<const scaled::FastElem *>(orbit_fast);
        scaled::CRed c = scaled::reduce_c(dcX, dcY);
        // keep the four words of c in registers (otherwise they are re-derived from dcX/dcY in every round)
        asm volatile("" : "+r"(c.xb), "+r"(c.yb), "+r"(c.xe), "+r"(c.ye));
        const unsigned lane_bit = 1u << (threadIdx.x & 31);
        scaled::Lane L;
        scaled::Mode mode = live ? scaled::kTry : scaled::kDone;
        for (;;) {
            if (mode == scaled::kTry)
                mode = scaled::enter<IterT>(tab, c, a.dxm, a.dxe, a.dym, a.dye, RefIteration, iter, n_iterations, L)
                           ? scaled::kFast : scaled::kSlow;
            const unsigned fast_m = __ballot_sync(0xffffffffu, mode == scaled::kFast);
            const unsigned slow_m = __ballot_sync(0xffffffffu, mode == scaled::kSlow);
            if ((fast_m | slow_m) == 0u) break;
            if (mode == scaled::kFast) {
                mode = scaled::fast_round<IterT, Count>(tab, last, n_iterations, c, L, RefIteration, iter, a.dxm, a.dxe, a.dym,
                                                        a.dye, steps);
            } else if ((slow_m & lane_bit) && (fast_m == 0u || __popc(slow_m) >= kSlowBatch)) {
                // one float+exponent step for the lanes the scaled form refused
                a.z = __ldg(orb + RefIteration);
                if (!step<IterT, Count>(a, b, orb, last, n_iterations, dcX, dcY, RefIteration, iter, steps)) mode = scaled::kDone;
                else { a = b; mode = scaled::kTry; }
            }
        }
        // the delta after the last step is not observable (only `iter` is written out)
    }
};

} // namespace fs
