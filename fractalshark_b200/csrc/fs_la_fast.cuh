// fs_la_fast.cuh -- one LAv2 step of the float+exponent (HDRx32) path, select-free.
//
// What it computes: exactly the step of GPU_LAReference::getLA + GPU_LAInfoDeep::Prepare/Evaluate + getZ
// (GPU_LAReference.h:271-303, GPU_LAInfoDeep.h:90-123, LAstep.h:157-185) on HDRFloatComplex<float> operands -- the
// same individually rounded binary32 operations in the same order, hence the same bits:
//     t     = 2*Ref + dz                      (complex add, shared-exponent alignment)
//     newdz = Reduce(dz * t)                  (complex mul as nvcc contracted it: fma(ar,br,-(ai*bi)), fma(ai,br,ar*bi))
//     unusable = cheb(newdz) >= LAThreshold   (lexicographic on (exponent, mantissa))
//     dz'   = newdz*ZCoeff + dc*CCoeff        z = Ref' + dz'
//     rebase-by-norm = Reduce(cheb(z)) < Reduce(cheb(dz'))
//
// How it differs from the reference-shaped code in fs_types.cuh (which stays the fallback): the three aligned additions
// scale BOTH operands -- one of the two multipliers is exactly 1 -- so there are no operand selects:
//     r = fma(b, 2^min(-d,0), a * 2^min(d,0)),   d = a.e - b.e,   multiplier fields = clamp(127 -/+ d, 0, 127)
// (one VIADDMNMX + one shift per multiplier).  `a * 2^min(d,0)` is exact while it stays normal, and when it does not
// the other operand is at least 2^100 times larger, so the sum rounds to the same value as the reference's single FMA.
// The multiplier drops to 0 at |d| >= 127 where the reference's HDRFloatComplex addition drops the operand at
// |d| >= 120 (HDRFloatComplex.h:219-247): a step that meets an exponent gap in [120, 127) is REFUSED.  So is any step
// that touches an exact complex zero (Reduce is a no-op there, HDRFloatComplex.h:473-527 / HDRFloat.h:414-456), a
// non-finite value, or a mantissa far outside [2^-60, 2^60] (where the exactness argument above could fail).  A refused
// step is recomputed by the caller with the reference-shaped operations; oracle/lockstep_check.cpp runs this header on
// the CPU beside the oracle's float+exponent step on whole frames and counts refusals and (zero) mismatches.
#pragma once
#include "fs_types.cuh"

namespace fs {
namespace lafast {

// clamp(127 + d, 0, 127): exponent field of 2^min(d, 0), zero once d <= -127
FS_HD int mul_field(int d) {
#ifdef __CUDA_ARCH__
    return __viaddmin_s32_relu(d, 127, 127);
#else
    const int v = d < 0 ? d + 127 : 127;
    return v < 0 ? 0 : v;
#endif
}
FS_HD float fabs_(float x) {
#ifdef __CUDA_ARCH__
    return fabsf(x);
#else
    return __builtin_fabsf(x);
#endif
}
// s is a finite positive number in [2^lo, 2^hi)
template <int LO, int HI> FS_HD bool in_range(float s) {
    return (f2u(s) - ((uint32_t)(LO + 127) << 23)) < (((uint32_t)(HI + 127) << 23) - ((uint32_t)(LO + 127) << 23));
}

struct C3 {
    float re, im;
    int e;
};

// a + b on shared-exponent complex numbers; `gap` collects "an exponent gap in [120, 127) was met"
FS_HD C3 cadd(float ar, float ai, int ae, float br, float bi, int be, bool &gap) {
    const int d = ae - be;
    const float ma = u2f((uint32_t)mul_field(d) << 23), mb = u2f((uint32_t)mul_field(-d) << 23);
    C3 r;
    r.re = fma_(br, mb, ar * ma);
    r.im = fma_(bi, mb, ai * ma);
    r.e = imax(ae, be);
    gap = gap || (uint32_t)(iabs(d) - EXP_DIFF_IGNORED) < 7u;
    return r;
}

struct StepOut {
    C3 dz;           // newdz*ZCoeff + dc*CCoeff
    C3 z;            // Ref' + dz
    bool unusable;   // Prepare refused the step (dz, z are not computed then)
    bool rebase;     // |z| < |dz| by Chebyshev norm
};

// One step.  Returns false when the step must be recomputed by the reference-shaped code (nothing is committed).
// refr/refi/refe: Ref of the current record; nr/ni/ne: Ref of the next record.
FS_HD bool step(float refr, float refi, int refe, float zcr, float zci, int zce, float ccr, float cci, int cce,
                float thm, int the, float nr, float ni, int ne, float dzr, float dzi, int dze, float dcr, float dci,
                int dce, StepOut &o) {
    bool gap = false;
    // Prepare  GPU_LAInfoDeep.h:90-105
    const C3 t = cadd(refr, refi, refe + 1, dzr, dzi, dze, gap);
    const float w_re = fma_(dzr, t.re, -(dzi * t.im));
    const float w_im = fma_(dzi, t.re, dzr * t.im);
    const float ws = fabs_(w_re) + fabs_(w_im);
    if (!in_range<-60, 60>(ws)) return false; // zero, NaN/Inf, or far from normalised
    const float wm = fabs_(w_re) > fabs_(w_im) ? fabs_(w_re) : fabs_(w_im);
    const int kf = (int)(f2u(wm) >> 23); // biased exponent of the larger part, in [66, 187]
    const float sc = u2f((uint32_t)(254 - kf) << 23);
    const float nr_ = w_re * sc, ni_ = w_im * sc;
    const int nwe = dze + t.e + kf - 127;
    // cheb(newdz) >= LAThreshold, lexicographic; the larger part's reduced mantissa is wm with its exponent field reset
    const float nm = u2f((f2u(wm) & 0x007fffffu) | 0x3f800000u);
    o.unusable = !(nwe < the || (nwe == the && nm < thm));
    if (o.unusable) return !gap;
    // Evaluate  GPU_LAInfoDeep.h:120-123
    const float p_re = fma_(nr_, zcr, -(ni_ * zci)), p_im = fma_(ni_, zcr, nr_ * zci);
    const float q_re = fma_(dcr, ccr, -(dci * cci)), q_im = fma_(dci, ccr, dcr * cci);
    o.dz = cadd(p_re, p_im, nwe + zce, q_re, q_im, dce + cce, gap);
    // getZ  LAstep.h:181-185
    o.z = cadd(nr, ni, ne, o.dz.re, o.dz.im, o.dz.e, gap);
    const float zs = fabs_(o.z.re) + fabs_(o.z.im), ds = fabs_(o.dz.re) + fabs_(o.dz.im);
    // both sums finite, non-zero and within [2^-60, 2^60): their product is a normal number in [2^-120, 2^120)
    if (gap || !in_range<-60, 60>(zs) || !in_range<-60, 60>(ds)) return false;
    const float zm = fabs_(o.z.re) > fabs_(o.z.im) ? fabs_(o.z.re) : fabs_(o.z.im);
    const float dm = fabs_(o.dz.re) > fabs_(o.dz.im) ? fabs_(o.dz.re) : fabs_(o.dz.im);
    // Reduce of the two norms (non-zero): exponent += biased - 127, mantissa field kept
    const int zne = o.z.e + (int)(f2u(zm) >> 23), dne = o.dz.e + (int)(f2u(dm) >> 23);
    const uint32_t zmb = f2u(zm) & 0x007fffffu, dmb = f2u(dm) & 0x007fffffu;
    o.rebase = zne < dne || (zne == dne && zmb < dmb);
    return true;
}

} // namespace lafast
} // namespace fs
