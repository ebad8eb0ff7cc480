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
// for which a step ARRIVING at Z_n can neither escape nor rebase.  th = 0 forces the exact tests (Z_n = 0, the
// last element, |Z_n| close to 2); th = NaN marks an element the plain-float step must not touch.
struct alignas(16) FastElem {
    float ax, ay, th, pad;
};

constexpr int kChunk = 8;         // speculative steps per commit
constexpr int kNorm = 16;         // on entry |w|_inf is normalised into [2^16, 2^17)
constexpr float kLo = 0x1p-6f;    // every component of every committed w must be >= kLo ...
constexpr float kHi = 0x1p34f;    // ... and |w|_inf < kHi at every chunk boundary (<= 2^62 inside a chunk)
constexpr int kCMaxExp = 30;      // c' must be below 2^31 on entry
constexpr int kCDropExp = -110;   // c' below 2^-109 is dropped (provably < half an ulp of every non-zero sum)
constexpr int kDeepK = -100;      // exact escape/rebase tests need 2^-k representable with headroom
constexpr int kZeroK = -80;       // a step on Z = 0 needs d itself in the normal range ...
constexpr int kZeroKc = -46;      // ... and, when c' was dropped, |d|^2 * 2^-k large enough to dominate it
constexpr int kMinZExp = -20;     // eligible orbit elements: both components in [2^-20, 4)

FS_HD float pow2i(int e) { return u2f((uint32_t)(e + 127) << 23); } // exact 2^e, e in [-126, 127]
FS_HD int fexp(float f) { return (int)((f2u(f) >> 23) & 0xffu) - 127; }
FS_HD bool is_nan(float f) { return f != f; }

// Table entry from one HDRx32 orbit element {x.m, x.e, y.e, y.m} (GPU_ReferenceIter.h:119-125).
FS_HD FastElem make_fast_elem(float xm, int xe, float ym, int ye, bool last) {
    FastElem e;
    e.ax = 0.0f; e.ay = 0.0f; e.th = 0.0f; e.pad = 0.0f;
    const bool zero = xm == 0.0f && ym == 0.0f;
    if (!zero) {
        const float axm = fabsf(xm), aym = fabsf(ym);
        const bool nrm = axm >= 0.25f && axm < 4.0f && aym >= 0.25f && aym < 4.0f;
        const int vx = xe + fexp(xm), vy = ye + fexp(ym); // floor(log2 |component|)
        const bool ok = nrm && vx >= kMinZExp && vy >= kMinZExp && vx <= 1 && vy <= 1;
        if (!ok) {
            e.th = u2f(0x7fc00000u);
        } else {
            // 2*Z exactly: |m| in [0.25,4), exponent in [-20,1]  =>  two exact scalings
            e.ax = (xm * pow2i(xe + 1 + 30)) * pow2i(-30);
            e.ay = (ym * pow2i(ye + 1 + 30)) * pow2i(-30);
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
    if (last && !is_nan(e.th)) e.th = 0.0f; // arriving at the last element always rebases (LAKernel.cuh:214)
    return e;
}

FS_HD FastElem load_elem(const FastElem *tab, uint64_t n) {
#ifdef __CUDA_ARCH__
    const float4 v = __ldg(reinterpret_cast<const float4 *>(tab) + n);
    FastElem e;
    e.ax = v.x; e.ay = v.y; e.th = v.z; e.pad = v.w;
    return e;
#else
    return tab[n];
#endif
}

// Per-pixel constants of the scaled form.
struct Scale {
    int k;
    float sk;   // 2^k        (0 when not representable: d is then negligible against every eligible Z)
    float sk2;  // 2^(k+1)
    float ik;   // 2^-k       (clamped to 2^126: only makes the threshold test more conservative)
    float ccx, ccy; // c * 2^-k (0 when dropped)
    bool ckept;     // c' is exact (nothing was dropped)
};

enum Outcome : int { kFinished = 0, kContinue = 1, kNeedSlow = 2 };

// Try to express the float+exponent state (dX, dY) at orbit index n in scaled form.
template <class IterT>
FS_HD bool enter(const FastElem *tab, Hdr<float> cX, Hdr<float> cY, float dxm, int dxe, float dym, int dye, IterT n,
                 IterT iter, IterT n_iterations, Scale &sc, float &wx, float &wy, FastElem &E0) {
    if (dxm == 0.0f || dym == 0.0f) return false;
    if ((uint64_t)(n_iterations - iter) < (uint64_t)kChunk) return false;
    const int vx = dxe + fexp(dxm), vy = dye + fexp(dym);
    const int emax = imax(vx, vy), emin = imin(vx, vy);
    const int k = emax - kNorm;
    if (emin - k < -5) return false;
    E0 = load_elem(tab, n);
    if (is_nan(E0.th)) return false;
    // mantissas are floats in (2^-24 .. 4): scale in two exact steps to stay inside the normal range
    wx = (dxm * pow2i(imin(dxe - k, 60))) ;
    wy = (dym * pow2i(imin(dye - k, 60))) ;
    if (dxe - k > 60 || dye - k > 60 || dxe - k < -60 || dye - k < -60) return false;
    sc.k = k;
    sc.sk = k >= -126 ? pow2i(k) : 0.0f;
    sc.sk2 = k + 1 >= -126 ? pow2i(k + 1) : 0.0f;
    sc.ik = pow2i(imin(-k, 126));
    // c' : reduce c, then scale
    float cxs = 0.0f, cys = 0.0f;
    bool kept = true;
    if (cX.m != 0.0f) {
        const int ex = cX.e + fexp(cX.m) - k;
        if (ex > kCMaxExp) return false;
        if (ex >= kCDropExp) cxs = u2f((f2u(cX.m) & 0x807fffffu) | ((uint32_t)(ex + 127) << 23));
        else kept = false;
    }
    if (cY.m != 0.0f) {
        const int ey = cY.e + fexp(cY.m) - k;
        if (ey > kCMaxExp) return false;
        if (ey >= kCDropExp) cys = u2f((f2u(cY.m) & 0x807fffffu) | ((uint32_t)(ey + 127) << 23));
        else kept = false;
    }
    sc.ccx = cxs; sc.ccy = cys; sc.ckept = kept;
    if (k > 0) return false; // |d| >= 2^16: about to escape, not worth it (and keeps 2^-k >= 1)
    if (E0.ax == 0.0f && E0.ay == 0.0f && !(k >= kZeroK && (k >= kZeroKc || kept))) return false;
    return true;
}

// Run speculative chunks from the scaled state until the pixel finishes, a chunk is rejected, the budget runs
// out or w needs renormalising.  On return (dxm, dxe, dym, dye, n, iter) hold the last committed state in
// float+exponent form (reduced mantissas).
template <class IterT, bool Count>
FS_HD Outcome run(const FastElem *tab, IterT last, IterT n_iterations, const Scale sc, float wx, float wy, FastElem E,
                  float &dxm, int &dxe, float &dym, int &dye, IterT &RefIteration, IterT &iter,
                  unsigned long long &steps) {
    IterT n = RefIteration;
    Outcome out = kContinue;
    for (;;) {
        // committed state at the chunk boundary
        const float wx0 = wx, wy0 = wy;
        const IterT n0 = n;
        float lo = 0x1p100f;
        bool ok = true, done = false;
        int s = 0;
#pragma unroll
        for (s = 0; s < kChunk; s++) {
            const FastElem En = load_elem(tab, (uint64_t)n + 1);
            const float Sx = fma_(wx, sc.sk, E.ax), Sy = fma_(wy, sc.sk, E.ay); // 2Z + d
            const float pa = wx * Sx, pb = wy * Sy, pc = wx * Sy, pd = wy * Sx;
            const float sumX = pa - pb, sumY = pc + pd;
            float nx = sumX + sc.ccx, ny = sumY + sc.ccy;
            ++n;
            FastElem Ec = En;
            const float m = fmaxf(fabsf(nx), fabsf(ny));
            lo = fminf(fminf(fabsf(nx), fabsf(ny)), lo);
            const float thr = En.th * sc.ik;
            if (!(m < thr)) {
                // ---- exact escape and rebase tests (LAKernel.cuh:196-226), in scaled plain floats ----
                if (is_nan(En.th) || sc.k < kDeepK) { ok = false; break; }
                const float tx = fma_(nx, sc.sk2, En.ax), ty = fma_(ny, sc.sk2, En.ay); // 2*(Z' + d')
                const float tx2 = tx * tx, ty2 = ty * ty;
                const float n2 = tx2 + ty2;
                if (!(n2 < 16.0f)) { done = true; break; }
                const float hik = 0.5f * sc.ik;
                const float Tx = fma_(En.ax, hik, nx), Ty = fma_(En.ay, hik, ny); // (Z' + d') * 2^-k
                const float Tx2 = Tx * Tx, Ty2 = Ty * Ty, dx2 = nx * nx, dy2 = ny * ny;
                const float N2 = Tx2 + Ty2, D2 = dx2 + dy2;
                if (N2 < D2 || n >= last) {
                    // the next step runs on Z_0 = 0
                    if (!(sc.k >= kZeroK && (sc.k >= kZeroKc || sc.ckept))) { ok = false; break; }
                    nx = Tx; ny = Ty;
                    n = 0;
                    Ec = load_elem(tab, 0);
                    lo = fminf(fminf(fabsf(nx), fabsf(ny)), lo);
                }
            }
            wx = nx; wy = ny; E = Ec;
        }
        const float mend = fmaxf(fabsf(wx), fabsf(wy));
        if (!(lo >= kLo)) { ok = false; done = false; }
        if (!ok) {
            // discard the chunk
            wx = wx0; wy = wy0; n = n0;
            out = kNeedSlow;
            break;
        }
        if (done) {
            iter += (IterT)s;
            if (Count) steps += (unsigned long long)s + 1;
            RefIteration = n;
            return kFinished;
        }
        iter += (IterT)kChunk;
        if (Count) steps += (unsigned long long)kChunk;
        if (!(mend < kHi) || (uint64_t)(n_iterations - iter) < (uint64_t)kChunk) { out = kContinue; break; }
    }
    // back to float+exponent form (components are non-zero: lo >= kLo on every committed chunk, entry checked)
    RefIteration = n;
    dxm = u2f((f2u(wx) & 0x807fffffu) | 0x3f800000u); dxe = sc.k + fexp(wx);
    dym = u2f((f2u(wy) & 0x807fffffu) | 0x3f800000u); dye = sc.k + fexp(wy);
    return out;
}

} // namespace scaled
} // namespace fs
