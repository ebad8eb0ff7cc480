// fs_at_fast.cuh -- the AT shortcut (ATInfo::PerformAT, ATInfo.h:155-188) as a mantissa recurrence, for the
// float+exponent types with a binary32 or binary64 mantissa.  FS_HD so that oracle/lockstep_check.cpp can run the very
// same functions on the CPU against the oracle's float+exponent restatement of the loop.
//
// The reference iterates  z <- z*z + c  on an HDRFloatComplex (two mantissas, one shared exponent) and leaves when
// Reduce(|z|^2) > SqrEscapeRadius.  c is reduced; when its exponent E is <= 0, z*z has exponent 2E <= E, so `add`
// (HDRFloatComplex.h:219-247) always lands on c's exponent: from the first pass on the exponent of z IS E, and
//     re' = fma(rr - ii, 2^E, c.re)          im' = fma(fma(re, im, re*im), 2^E, c.im)
// with rr = re*re, ii = im*im are exactly the operations the reference performs on the mantissas (for E == 0 it adds
// c to z*z instead of z*z to c: the same sum, one rounding).  Its escape test compares (exponent, mantissa) pairs
// of reduced positive numbers, i.e. values:  rr + ii > R.m * 2^(R.e - 2E).  NaN/Inf mantissas leave the loop here as
// they do there (their exponent field reduces to +128 / +1024).
// E > 0 (|c| >= 2: pixels far from the centre of a deep view, whose c = RefC + dc*CCoeff is large, or a view whose
// AT constant itself sits past 2 like View 5's |RefC| ~ 2.008) is left to the general loop: there the reference's
// exponent doubles every pass and its mantissas run into the denormal range, a behaviour only that loop reproduces.
#pragma once
#include "fs_types.cuh"

namespace fs {
namespace atfast {

template <class M> struct Plan {
    bool ok; // false: take the general float+exponent loop
    M s;     // 2^E
    M thr;   // R.m * 2^(R.e - 2E)
    int E;
    bool mono; // R > 4 and |c| <= R/4: past the escape radius |z| only grows (see lav2_at), so an escape anywhere in a
               // chunk of passes is still visible in the chunk's last pass
};

// c: reduced; R = SqrEscapeRadius.  `passes` = n_iterations / StepLength (the plan needs at least one pass: the
// first pass is what moves z from its initial MIN_BIG exponent to E).
template <class M> FS_HD Plan<M> plan(HdrC<M> c, Hdr<M> R, bool passes) {
    Plan<M> p;
    constexpr int kMaxShift = MT<M>::BIAS - 1;
    const int sh = R.e - 2 * c.e;
    p.E = c.e;
    p.ok = passes && c.e <= 0 && c.e > -EXP_DIFF_IGNORED && R.m >= M(1) && R.m < M(2) && sh <= kMaxShift && sh >= -kMaxShift;
    p.s = p.ok ? MT<M>::pow2(c.e) : M(0);
    p.thr = p.ok ? R.m * MT<M>::pow2(sh) : M(0);
    // real-valued R and |c|^2 (R <= 2^32 for the 32-bit exponent range the table builder uses, LAInfoDeep.h:486-494;
    // anything that overflows or is NaN fails the comparisons and takes the running-maximum form)
    const M Rv = R.m * MT<M>::pow2(R.e);
    const M c2 = (c.re * c.re + c.im * c.im) * p.s * p.s;
    p.mono = p.ok && R.e < 40 && Rv > M(4) && c2 <= Rv * Rv * M(0.0625);
    return p;
}

// |z|^2 mantissa of the current z
template <class M> FS_HD M norm(M re, M im) { return re * re + im * im; }
template <class M> FS_HD bool escaped(M nsq, M thr) { return !(nsq <= thr); }
// The reference's imaginary part of z*z is  t = fma(re, im, RN(re*im))  (nvcc's contraction of re*im + im*re).  With
// p = RN(re*im) and x = re*im exact:  x + p = 2p - (p - x), and |p - x| <= ulp(p)/2 = ulp(2p)/4 (at a binade edge the
// lower neighbour of 2p is ulp(2p)/2 away, still twice the error; in the denormal range |p - x| <= 2^-150 is exactly
// half the spacing of 2p's neighbours and the tie goes to the even one, 2p), so  t == 2p  for every input -- overflow
// included, where both are +-inf, and NaN.  Hence  fma(t, s, c.im) == fma(p, 2s, c.im)  whenever 2p is finite (a power of
// two moved between the factors of an exact product), and a non-finite 2p needs |re*im| > FLT_MAX/2, i.e. |z|^2 >=
// 2|re*im| overflowed: the escape test ahead of that pass has already left the loop (lav2_at replays such a chunk).
// One rounding and one FMA-pipe slot less per pass: 6 instead of 7.
template <class M> FS_HD void advance(M &re, M &im, M s, M cre, M cim) {
    const M rr = re * re, ii = im * im;
    const M p = re * im;
    re = fma_(rr - ii, s, cre);
    im = fma_(p, s + s, cim);
}
// the reference-shaped pass, kept for the checker (oracle/lockstep_check.cpp compares the two forms pass by pass)
template <class M> FS_HD void advance_as_written(M &re, M &im, M s, M cre, M cim) {
    const M rr = re * re, ii = im * im;
    const M t = fma_(re, im, re * im);
    re = fma_(rr - ii, s, cre);
    im = fma_(t, s, cim);
}

} // namespace atfast

// Passes per escape test of the chunked AT loop, and per comparison of the cycle watch (fs_lav2.cuh lav2_at).
#ifndef FS_AT_CHUNK
#define FS_AT_CHUNK 16
#endif
constexpr int kWatchChunk = FS_AT_CHUNK;
FS_HD int first_bit(unsigned int m) { // 1-based index of the lowest set bit
#ifdef __CUDA_ARCH__
    return __ffs((int)m);
#else
    return __builtin_ffs((int)m);
#endif
}
// Cycle watch of the chunked AT loop: see the comment ahead of lav2_at in fs_lav2.cuh.  FS_HD so that
// oracle/lockstep_check.cpp runs the very same code against the loop that executes every pass.
template <class M, class IterT> struct CycleWatch {
    M sre, sim;
    IterT at;       // pass count of the saved state
    IterT next;     // chunk-boundary pass count at which the next state is saved
    bool armed;
    FS_HD CycleWatch(M re, M im, IterT i) : sre(re), sim(im), at(i), next(i + (IterT)kWatchChunk), armed(true) {}
    // `seen`: bit u set = the state after pass u + 1 of the chunk just finished (which ended at pass count i) equalled the
    // saved state.  Any hit gives a period (a multiple of the true one): the lowest bit the shortest.
    FS_HD void after_chunk_seen(M re, M im, IterT &i, IterT at_max, IterT &skipped, unsigned int seen) {
        if (!armed) return;
        if (seen != 0u) {
            const IterT hit = i - (IterT)kWatchChunk + (IterT)first_bit(seen); // pass count at the first hit
            const IterT P = hit - at;
            if (P != 0) {
                skipped = ((at_max - i) / P) * P;
                i += skipped;
                armed = false;
                return;
            }
        }
        if (i == next) {
            sre = re; sim = im;
            next = i + (i - at) * 2;
            at = i;
        }
    }
    FS_HD void after_chunk(M re, M im, IterT &i, IterT at_max, IterT &skipped) {
        if (!armed) return;
        if (bits_equal(re, sre) && bits_equal(im, sim)) {
            const IterT P = i - at;
            skipped = ((at_max - i) / P) * P;
            i += skipped;
            armed = false;
        } else if (i == next) {
            sre = re; sim = im;
            next = i + (i - at) * 2; // gaps of 1, 2, 4, 8 ... chunks (stops growing if it would wrap: i never gets there)
            at = i;
        }
    }
};


} // namespace fs
