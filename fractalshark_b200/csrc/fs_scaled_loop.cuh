// fs_scaled_loop.cuh -- HDRx32 perturbation steps evaluated in plain binary32 on a per-pixel scaled delta.
//
// What it computes: exactly the float+exponent step of LAKernel.cuh:130-236 / HDRFloat.h:725-794 (the same
// sequence of individually rounded operations), hence bit-identical iteration counts.
//
// Why it can be done in plain floats: every HDRFloat<float> operation rounds a 24-bit mantissa and keeps the
// exponent in a separate integer, i.e. it is binary32 arithmetic with an unbounded exponent.  A power-of-two
// rescaling commutes with every rounding as long as no intermediate leaves the normal binary32 range, so with
//     d = w * 2^k   (w = (wx, wy) plain floats, k one integer per pixel, c' = c * 2^-k)
// the step   S = 2Z + d ; w' = (wx*Sx - wy*Sy + c'x , wx*Sy + wy*Sx + c'y)   is 2 FFMA + 4 FMUL + 4 FADD with
// no exponent work at all (the float+exponent form needs ~140 instructions for the same 20 roundings, see
// profiles/r1_lav2_v2_summary.md).  The escape test |Z'+d'|^2 < 4 and the rebase test |Z'+d'|^2 < |d'|^2 cannot
// fire while |d'| is small against Z'; one precomputed threshold per orbit element (`th`) decides that with
// one FMNMX + FMUL + FSETP, and only steps past the threshold evaluate the two tests (exactly, also in plain
// floats).
//
// Validity is checked, not assumed: a chunk of kChunk steps runs speculatively from a saved state and commits
// only if every intermediate provably stayed in range (components of w within [2^-6, 2^62], orbit elements
// flagged eligible when the table is built, c' either exactly representable or provably below half an ulp of
// every sum it is added to).  Otherwise the chunk is discarded and the pixel takes one float+exponent step
// (fs_perturb_loop.cuh, hdr32fast::step) before trying again.  Exact zeros, which the reference treats with
// non-canonical exponents (HDRFloat.h:974-1000 vs :725-794), always go to that path.
//
// The arithmetic in this header is FS_HD so tests/lockstep_check.cpp can run it on the CPU in lockstep with
// the oracle's float+exponent restatement (every step compared by value).
#pragma once
#include <math.h>
#include "fs_types.cuh"

namespace fs {
namespace scaled {

// One entry per reference-orbit element n:  ax, ay = 2*Z_n as plain floats;  th = largest |d|_inf (true scale)
// for which a step ARRIVING at Z_n can neither escape nor rebase;  kmin = smallest scale exponent k for which
// the plain-float step is exact on this element (as a float; checked only on the th = 0 classes).
//   regular element (both components in [2^-20, 4)):  th > 0,  kmin = -inf
//   Z_n = 0, tiny components (down to 2^-77), the last element, |Z_n| close to 2:  th = 0 (always run the exact tests)
//   anything else (one component zero, out of range, unnormalised mantissa):  th = NaN (never touched)
struct alignas(16) FastElem {
    float ax, ay, th, kmin;
};

constexpr int kChunk = 16;        // speculative steps per commit
constexpr int kNorm = 16;         // (re)normalisation puts |w|_inf into [2^16, 2^17)
constexpr float kLo = 0x1p-6f;    // every component of every committed w must be >= kLo ...
constexpr float kHi = 0x1p40f;    // ... and |w|_inf < kHi at every chunk boundary (< 2^95 inside a chunk: <= x10 per step)
constexpr float kTestScale = 0x1p-40f; // the rebase test squares w * 2^-40 (keeps (2^95)^2 inside binary32; a square
                                       // that underflows is > 2^20 below the other side, the decision is unaffected)
constexpr int kCMaxExp = 30;      // c' must be below 2^31
constexpr int kCDropExp = -110;   // c' below 2^-109 is dropped: any sum it could change is rejected by kLo anyway
constexpr int kDeepK = -126;      // the exact escape/rebase tests need 2^k and 2^-k representable
constexpr int kNegligibleK = -200; // below this, d is < 2^-25 of every eligible non-zero orbit component (>= 2^-77)
constexpr float kZeroKmin = -100.0f; // a step on Z = 0 needs d itself well inside the normal range
constexpr float kTinyKmin = -126.0f; // elements with a component below 2^-20 need 2^k exact (or d negligible, see chunk)
constexpr int kMinZExp = -20, kTinyZExp = -77;

FS_HD float pow2i(int e) { return u2f((uint32_t)(e + 127) << 23); } // exact 2^e, e in [-126, 127]
FS_HD int fexp(float f) { return (int)((f2u(f) >> 23) & 0xffu) - 127; }
FS_HD bool is_nan(float f) { return f != f; }

// Table entry from one HDRx32 orbit element {x.m, x.e, y.e, y.m} (GPU_ReferenceIter.h:119-125).
FS_HD FastElem make_fast_elem(float xm, int xe, float ym, int ye, bool last) {
    FastElem e;
    e.ax = 0.0f; e.ay = 0.0f; e.th = 0.0f; e.kmin = kZeroKmin;
    const bool zero = xm == 0.0f && ym == 0.0f;
    if (!zero) {
        const float axm = fabsf(xm), aym = fabsf(ym);
        const bool nrm = axm >= 0.25f && axm < 4.0f && aym >= 0.25f && aym < 4.0f;
        const int vx = xe + fexp(xm), vy = ye + fexp(ym); // floor(log2 |component|)
        const bool ok = nrm && vx >= kTinyZExp && vy >= kTinyZExp && vx <= 1 && vy <= 1;
        if (!ok) {
            e.th = u2f(0x7fc00000u);
        } else {
            // 2*Z exactly: |m| in [0.25,4), exponent in [-77,1]  =>  two exact scalings
            e.ax = (xm * pow2i(xe + 1 + 60)) * pow2i(-60);
            e.ay = (ym * pow2i(ye + 1 + 60)) * pow2i(-60);
            if (vx < kMinZExp || vy < kMinZExp) {
                // th stays 0.  |Z|_inf < 1 keeps |Z|^2 < 2: with d negligible such an element cannot escape
                if (vx >= 0 || vy >= 0) e.th = u2f(0x7fc00000u);
                e.kmin = kTinyKmin;
            } else {
                e.kmin = -3.0e38f;
                const double a = 0.5 * fabs((double)e.ax), b = 0.5 * fabs((double)e.ay);
                const double margin = 1.0 - 0x1p-18;
                // no rebase while |d|_2 < |Z|_2 / 2, and |d|_2 <= sqrt(2) |d|_inf
                const double th_rebase = sqrt(a * a + b * b) * 0.35355339059327373 * margin;
                // no escape while (a+g)^2 + (b+g)^2 < 4 with g = |d|_inf
                const double s = a + b, q = a * a + b * b, lim = 4.0 * margin;
                const double disc = s * s - 2.0 * (q - lim);
                double g = disc > 0.0 ? 0.5 * (sqrt(disc) - s) * margin : 0.0;
                if (!(g > 0.0)) g = 0.0;
                e.th = (float)(th_rebase < g ? th_rebase : g);
            }
        }
    }
    if (last && !is_nan(e.th)) e.th = 0.0f; // arriving at the last element always rebases (LAKernel.cuh:214)
    return e;
}

FS_HD FastElem load_elem(const FastElem *tab, uint64_t n) {
#ifdef __CUDA_ARCH__
    const float4 v = __ldg(reinterpret_cast<const float4 *>(tab) + n);
    FastElem e;
    e.ax = v.x; e.ay = v.y; e.th = v.z; e.kmin = v.w;
    return e;
#else
    return tab[n];
#endif
}

// c = (cX, cY) of the pixel with reduced mantissas: sign|mantissa bits and value exponent per component
// (a zero component gets an exponent far below every drop threshold).
struct CRed {
    uint32_t xb, yb;
    int xe, ye;
};
FS_HD CRed reduce_c(Hdr<float> cX, Hdr<float> cY) {
    CRed c;
    c.xb = f2u(cX.m) & 0x807fffffu; c.xe = cX.m != 0.0f ? cX.e + fexp(cX.m) : -(1 << 30);
    c.yb = f2u(cY.m) & 0x807fffffu; c.ye = cY.m != 0.0f ? cY.e + fexp(cY.m) : -(1 << 30);
    return c;
}

// Per-pixel constants of the scaled form  d = w * 2^k.
struct Scale {
    int k;
    float sk;   // 2^k        (0 when not representable: d is then negligible against every eligible Z)
    float sk2;  // 2^(k+1)
    float ik;   // 2^-k       (clamped to 2^126: only makes the threshold test more conservative)
    float ccx, ccy; // c * 2^-k (0 when dropped)
};
FS_HD bool set_scale(Scale &sc, int k, const CRed &c) {
    const int ex = c.xe - k, ey = c.ye - k;
    if (k > 0 || ex > kCMaxExp || ey > kCMaxExp) return false;
    sc.k = k;
    sc.sk = k >= -126 ? pow2i(k) : 0.0f;
    sc.sk2 = k >= -127 ? pow2i(k + 1) : 0.0f;
    sc.ik = pow2i(imin(-k, 126));
    sc.ccx = ex >= kCDropExp ? u2f(c.xb | ((uint32_t)(ex + 127) << 23)) : 0.0f;
    sc.ccy = ey >= kCDropExp ? u2f(c.yb | ((uint32_t)(ey + 127) << 23)) : 0.0f;
    return true;
}
// The element a step starts from must allow the current scale.
FS_HD bool elem_allows(const FastElem &E, int k) {
    return !is_nan(E.th) && ((float)k >= E.kmin || (k <= kNegligibleK && E.kmin == kTinyKmin));
}

// Try to express the float+exponent state (dX, dY) at orbit index n in scaled form.
template <class IterT>
FS_HD bool enter(const FastElem *tab, const CRed &c, float dxm, int dxe, float dym, int dye, IterT n, IterT iter,
                 IterT n_iterations, Scale &sc, float &wx, float &wy, FastElem &E0) {
    if (!(fabsf(dxm) >= 0x1p-60f && fabsf(dym) >= 0x1p-60f)) return false; // zero, NaN or far from normalised
    if ((uint64_t)(n_iterations - iter) < (uint64_t)kChunk) return false;
    const int vx = dxe + fexp(dxm), vy = dye + fexp(dym);
    const int k = imax(vx, vy) - kNorm;
    if (imin(vx, vy) - k < -5) return false;
    if (!set_scale(sc, k, c)) return false;
    E0 = load_elem(tab, n);
    if (!elem_allows(E0, k)) return false;
    // mantissas are floats in (2^-24, 4) and the values land in [2^-5, 2^17): one exact scaling each
    wx = dxm * pow2i(dxe - k);
    wy = dym * pow2i(dye - k);
    return true;
}

// Re-centre w on 2^kNorm without leaving the scaled form (exact: a power-of-two shift of w, k and c').
// Returns false (state untouched) when the shifted state would violate an entry condition.
FS_HD bool renorm(Scale &sc, const CRed &c, float &wx, float &wy, const FastElem &E) {
    const float m = fmaxf(fabsf(wx), fabsf(wy));
    const int e = fexp(m) - kNorm;
    if (e == 0) return true;
    if (e > 100 || e < -100) return false;
    const float f = pow2i(-e);
    const float nwx = wx * f, nwy = wy * f;
    if (!(fminf(fabsf(nwx), fabsf(nwy)) >= 0x1p-5f)) return false;
    Scale ns;
    if (!set_scale(ns, sc.k + e, c) || !elem_allows(E, sc.k + e)) return false;
    sc = ns; wx = nwx; wy = nwy;
    return true;
}

enum ChunkResult : int { kCommitted = 0, kFinished = 1, kRejected = 2 };

FS_HD FastElem load_elem_at(const FastElem *p, int off) {
#ifdef __CUDA_ARCH__
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + off);
    FastElem e;
    e.ax = v.x; e.ay = v.y; e.th = v.z; e.kmin = v.w;
    return e;
#else
    return p[off];
#endif
}

// One speculative chunk: up to kChunk steps from (wx, wy, E = element at n), straight-line; the first step whose
// result passes the element's threshold ends the chunk and gets the exact escape / rebase tests (one shared
// copy of that code).  kCommitted: state advanced by `steps` iterations.  kFinished: the pixel escaped after
// `steps` further iterations (steps + 1 executed).  kRejected: state untouched, steps = 0.
template <class IterT>
FS_HD ChunkResult chunk(const FastElem *tab, IterT last, const Scale &sc, float &wx, float &wy, FastElem &E, IterT &n,
                        int &steps) {
    const float wx0 = wx, wy0 = wy;
    const IterT n0 = n;
    const FastElem *p = tab + n0; // step u arrives at p[u + 1]
    float lo = 0x1p100f;
    int cnt = kChunk;
    bool trig = false;
#pragma unroll
    for (int u = 0; u < kChunk; u++) {
        const FastElem En = load_elem_at(p, u + 1);
        const float Sx = fma_(wx, sc.sk, E.ax), Sy = fma_(wy, sc.sk, E.ay); // 2Z + d
        const float pa = wx * Sx, pb = wy * Sy, pc = wx * Sy, pd = wy * Sx;
        const float sumX = pa - pb, sumY = pc + pd;
        wx = sumX + sc.ccx;
        wy = sumY + sc.ccy;
        E = En;
        const float m = fmaxf(fabsf(wx), fabsf(wy));
        lo = fminf(fminf(fabsf(wx), fabsf(wy)), lo);
        if (!(m < En.th * sc.ik)) { cnt = u + 1; trig = true; break; }
    }
    n = n0 + (IterT)cnt;
    bool ok = lo >= kLo, done = false;
    if (trig && ok) {
        // ---- the step that arrived at E = element n: exact escape and rebase tests (LAKernel.cuh:196-226) ----
        if (!elem_allows(E, sc.k)) {
            ok = false;
        } else if (sc.k < kDeepK) {
            // d is negligible against Z' (both components in [2^-77, 1)): no escape, no rebase by norm;
            // everything else at this depth takes the float+exponent step
            if (sc.k > kNegligibleK || E.kmin != kTinyKmin || n >= last) ok = false;
        } else {
            const float tx = fma_(wx, sc.sk2, E.ax), ty = fma_(wy, sc.sk2, E.ay); // 2*(Z' + d')
            const float tx2 = tx * tx, ty2 = ty * ty;
            const float n2 = tx2 + ty2;
            if (!(n2 < 16.0f)) {
                done = true;
            } else {
                const float Tx = fma_(E.ax, 0.5f * sc.ik, wx), Ty = fma_(E.ay, 0.5f * sc.ik, wy); // (Z' + d') * 2^-k
                const float Txs = Tx * kTestScale, Tys = Ty * kTestScale, dxs = wx * kTestScale, dys = wy * kTestScale;
                const float Tx2 = Txs * Txs, Ty2 = Tys * Tys, dx2 = dxs * dxs, dy2 = dys * dys;
                const float N2 = Tx2 + Ty2, D2 = dx2 + dy2;
                if (N2 < D2 || n >= last) {
                    // rebase: the next step runs on Z_0
                    E = load_elem(tab, 0);
                    ok = elem_allows(E, sc.k) && fmaxf(fabsf(Tx), fabsf(Ty)) < kHi && fminf(fabsf(Tx), fabsf(Ty)) >= kLo;
                    wx = Tx; wy = Ty;
                    n = 0;
                }
            }
        }
    }
    if (!ok) {
        wx = wx0; wy = wy0; n = n0;
        E = load_elem(tab, n0);
        steps = 0;
        return kRejected;
    }
    if (done) { steps = cnt - 1; return kFinished; }
    steps = cnt;
    return kCommitted;
}

// Back to float+exponent form (reduced mantissas; components are non-zero by construction).
FS_HD void leave(const Scale &sc, float wx, float wy, float &dxm, int &dxe, float &dym, int &dye) {
    dxm = u2f((f2u(wx) & 0x807fffffu) | 0x3f800000u); dxe = sc.k + fexp(wx);
    dym = u2f((f2u(wy) & 0x807fffffu) | 0x3f800000u); dye = sc.k + fexp(wy);
}

// ---- per-pixel control flow shared by the kernel (fs_perturb_loop.cuh) and the CPU lockstep checker ----------
enum Mode : int { kDone = 0, kFast = 1, kSlow = 2, kTry = 3 };

struct Lane {
    Scale sc;
    float wx, wy;
    FastElem E; // table entry of the orbit element the next step starts from
};

// One fast-mode iteration of one pixel: optional re-centring, one speculative chunk, bookkeeping.
// Returns the pixel's next mode; whenever that is kSlow, (dxm, dxe, dym, dye) hold the committed state in
// float+exponent form.  n is the orbit index (RefIteration) in either form.
template <class IterT, bool Count>
FS_HD Mode fast_iteration(const FastElem *tab, IterT last, IterT n_iterations, const CRed &c, bool recenter, Lane &L,
                          IterT &n, IterT &iter, float &dxm, int &dxe, float &dym, int &dye,
                          unsigned long long &steps) {
    if (recenter && !renorm(L.sc, c, L.wx, L.wy, L.E)) {
        leave(L.sc, L.wx, L.wy, dxm, dxe, dym, dye);
        return kSlow;
    }
    int cnt = 0;
    const ChunkResult r = chunk<IterT>(tab, last, L.sc, L.wx, L.wy, L.E, n, cnt);
    if (r == kFinished) {
        iter += (IterT)cnt;
        if (Count) steps += (unsigned long long)cnt + 1;
        return kDone;
    }
    if (r == kCommitted) {
        iter += (IterT)cnt;
        if (Count) steps += (unsigned long long)cnt;
        const bool budget = (uint64_t)(n_iterations - iter) >= (uint64_t)kChunk;
        if (budget && fmaxf(fabsf(L.wx), fabsf(L.wy)) < kHi) return kFast;
        if (budget && renorm(L.sc, c, L.wx, L.wy, L.E)) return kFast;
    }
    leave(L.sc, L.wx, L.wy, dxm, dxe, dym, dye);
    return kSlow;
}

} // namespace scaled
} // namespace fs
