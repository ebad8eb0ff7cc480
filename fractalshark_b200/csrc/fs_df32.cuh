// fs_df32.cuh -- 2x32 "double-float" arithmetic and the two numeric policies built on it:
//   Num2x32     <- T = CudaDblflt<MattDblflt>            (Gpu2x32Perturbed*, Gpu2x32)
//   NumHdr2x32  <- T = HDRFloat<CudaDblflt<MattDblflt>>  (GpuHDRx2x32Perturbed*, GpuHDRx32 direct)
//
// Behavioural contract: HpSharkFloatLib/dblflt.cuh:68-317 (error-free transformations after Thall / Nagai et
// al.: every step is an individually rounded binary32 add, multiply or FMA), CudaDblflt.h:25-282 (operators,
// lexicographic comparisons, abs), HDRFloat.h:414-551, 829-1065 (float+exponent wrapper with a 2x32 mantissa)
// and HDRFloatComplex.h:158-173, 219-348, 473-527 / FloatComplex.h (complex forms).  The reference spells
// every one of these operations with __fadd_rn/__fmul_rn/__fmaf_rn, which the compiler may neither contract
// nor reassociate, so the value of each operation is fixed by its operation sequence; the sequences below are
// the published algorithms with the same operand order.
// Reproduced quirks (not "fixed"): `a <= b` on the 2x32 type evaluates `!(b > a)`, i.e. a >= b
// (CudaDblflt.h:221-225); multiplying by the constant 2 goes through the full product + renormalisation.
#pragma once
#include "fs_types.cuh"

namespace fs {

// ---- df32 primitives --------------------------------------------------------------------------------------
FS_D df32 df_make(float h, float t) { df32 z; z.head = h; z.tail = t; return z; }
// MattDblflt(float a, float b)  dblflt.h:19-28 (plain operators; nothing here can be contracted)
FS_D df32 df_two_sum(float a, float b) {
    df32 z;
    z.head = __fadd_rn(a, b);
    float t1 = __fadd_rn(z.head, -a);
    float t2 = __fadd_rn(z.head, -t1);
    t1 = __fadd_rn(b, -t1);
    t2 = __fadd_rn(a, -t2);
    z.tail = __fadd_rn(t1, t2);
    return z;
}
FS_D df32 df_from_float(float a) { return df_two_sum(a, 0.0f); } // explicit MattDblflt(float)  dblflt.h:54-55
FS_D df32 df_neg(df32 a) { return df_make(-a.head, -a.tail); }
// add_dblflt  dblflt.cuh:121-137
FS_D df32 df_add(df32 a, df32 b) {
    float t1 = __fadd_rn(a.head, b.head);
    float t2 = __fadd_rn(t1, -a.head);
    float t3 = __fadd_rn(__fadd_rn(a.head, __fadd_rn(t2, -t1)), __fadd_rn(b.head, -t2));
    float t4 = __fadd_rn(a.tail, b.tail);
    t2 = __fadd_rn(t4, -a.tail);
    const float t5 = __fadd_rn(__fadd_rn(a.tail, __fadd_rn(t2, -t4)), __fadd_rn(b.tail, -t2));
    t3 = __fadd_rn(t3, t4);
    t4 = __fadd_rn(t1, t3);
    t3 = __fadd_rn(__fadd_rn(t1, -t4), t3);
    t3 = __fadd_rn(t3, t5);
    df32 z;
    z.head = __fadd_rn(t4, t3);
    z.tail = __fadd_rn(__fadd_rn(t4, -z.head), t3);
    return z;
}
// sub_dblflt  dblflt.cuh:146-162
FS_D df32 df_sub(df32 a, df32 b) {
    float t1 = __fadd_rn(a.head, -b.head);
    float t2 = __fadd_rn(t1, -a.head);
    float t3 = __fadd_rn(__fadd_rn(a.head, __fadd_rn(t2, -t1)), -__fadd_rn(b.head, t2));
    float t4 = __fadd_rn(a.tail, -b.tail);
    t2 = __fadd_rn(t4, -a.tail);
    const float t5 = __fadd_rn(__fadd_rn(a.tail, __fadd_rn(t2, -t4)), -__fadd_rn(b.tail, t2));
    t3 = __fadd_rn(t3, t4);
    t4 = __fadd_rn(t1, t3);
    t3 = __fadd_rn(__fadd_rn(t1, -t4), t3);
    t3 = __fadd_rn(t3, t5);
    df32 z;
    z.head = __fadd_rn(t4, t3);
    z.tail = __fadd_rn(__fadd_rn(t4, -z.head), t3);
    return z;
}
// mul_dblflt  dblflt.cuh:169-180
FS_D df32 df_mul(df32 a, df32 b) {
    const float th = __fmul_rn(a.head, b.head);
    float tt = __fmaf_rn(a.head, b.head, -th);
    tt = __fmaf_rn(a.tail, b.tail, tt);
    tt = __fmaf_rn(a.head, b.tail, tt);
    tt = __fmaf_rn(a.tail, b.head, tt);
    df32 z;
    z.head = __fadd_rn(th, tt);
    z.tail = __fadd_rn(__fadd_rn(th, -z.head), tt);
    return z;
}
// sqr_dblflt  dblflt.cuh:198-219
FS_D df32 df_sqr(df32 a) {
    const float th = __fmul_rn(a.head, a.head);
    float tt = __fmaf_rn(a.head, a.head, -th);
    tt = __fmaf_rn(a.tail, a.tail, tt);
    const float e = __fmul_rn(a.head, a.tail);
    tt = __fmaf_rn(2.0f, e, tt);
    df32 z;
    z.head = __fadd_rn(th, tt);
    z.tail = __fadd_rn(__fadd_rn(th, -z.head), tt);
    return z;
}
// comparisons  CudaDblflt.h:200-247
FS_D bool df_lt(df32 a, df32 b) { return a.head < b.head || (a.head == b.head && a.tail < b.tail); }
FS_D bool df_eq(df32 a, df32 b) { return a.head == b.head && a.tail == b.tail; }
FS_D bool df_gt(df32 a, df32 b) { return !df_lt(a, b) && !df_eq(b, a); }
FS_D bool df_ge(df32 a, df32 b) { return !df_lt(a, b); }
FS_D bool df_le_as_written(df32 a, df32 b) { return !df_gt(b, a); } // CudaDblflt.h:221-225 (evaluates a >= b)
FS_D bool df_is_zero(df32 a) { return a.head == 0.0f && a.tail == 0.0f; }
// abs()  CudaDblflt.h:249-257
FS_D df32 df_abs(df32 a) { return df_lt(a, df_from_float(0.0f)) ? df_neg(a) : a; }
// (T)scalbnf(1, s) with the getMultiplier / getMultiplierNeg cut-offs  HDRFloat.h:497-551
FS_D df32 df_pow2(int s) { // getMultiplier: 0 at s <= -127; numeric_limits<CudaDblflt>::max() is value-initialised (0) at s >= 128
    if (s <= -127 || s >= 128) return df_make(0.0f, 0.0f);
    return df_from_float(__uint_as_float((uint32_t)(s + 127) << 23));
}
FS_D df32 df_pow2neg(int s) { // getMultiplierNeg (callers pass s <= 0)
    if (s <= -127) return df_make(0.0f, 0.0f);
    return df_from_float(scalbnf(1.0f, s));
}

// ---- plain scalar vocabulary for df32 (Num2x32::Real) -------------------------------------------------------
FS_D df32 add(df32 a, df32 b) { return df_add(a, b); }
FS_D df32 sub(df32 a, df32 b) { return df_sub(a, b); }
FS_D df32 mul(df32 a, df32 b) { return df_mul(a, b); }
FS_D void reduce(df32 &) {}
FS_D bool lt_pr(df32 a, df32 b) { return df_lt(a, b); }
FS_D bool ge_pr(df32 a, df32 b) { return df_ge(a, b); }
FS_D bool gt_pr(df32 a, df32 b) { return df_gt(a, b); }
FS_D bool le_pr(df32 a, df32 b) { return df_le_as_written(a, b); }
FS_D bool lt_bailout(df32 a) { return df_lt(a, df_from_float(256.0f)); } // one < T(256)  HDRFloat.h:1562-1586

// ---- FloatComplex<CudaDblflt>  (FloatComplex.h) --------------------------------------------------------------
FS_D Cx<df32> add(Cx<df32> a, Cx<df32> b) { Cx<df32> r; r.re = df_add(a.re, b.re); r.im = df_add(a.im, b.im); return r; }
FS_D Cx<df32> mul(Cx<df32> a, Cx<df32> b) {
    Cx<df32> r;
    r.re = df_sub(df_mul(a.re, b.re), df_mul(a.im, b.im));
    r.im = df_add(df_mul(a.re, b.im), df_mul(a.im, b.re));
    return r;
}
FS_D Cx<df32> mul(Cx<df32> a, df32 f) { Cx<df32> r; r.re = df_mul(a.re, f); r.im = df_mul(a.im, f); return r; }
FS_D void reduce(Cx<df32> &) {}
FS_D df32 cheb(Cx<df32> a) {
    const df32 ar = df_abs(a.re), ai = df_abs(a.im);
    return df_gt(ar, ai) ? ar : ai;
}
FS_D df32 norm2(Cx<df32> a) { return df_add(df_mul(a.re, a.re), df_mul(a.im, a.im)); }

// ---- HDRFloat<CudaDblflt>  (HDRFloat.h) ----------------------------------------------------------------------
FS_D Hdr<df32> hd_make(int e, df32 m) { Hdr<df32> r; r.m = m; r.e = e; return r; }
FS_D Hdr<df32> hd_zero() { return hd_make(MIN_BIG, df_make(0.0f, 0.0f)); }
// Reduce  HDRFloat.h:458-485: head exponent moves into exp; the tail's exponent field is rebased by the same
// amount, saturating at the zero field
FS_D void reduce(Hdr<df32> &a) {
    if (df_is_zero(a.m)) return;
    const uint32_t by = __float_as_uint(a.m.head), bx = __float_as_uint(a.m.tail);
    const int f_exp_y = (int)((by & 0x7F800000u) >> 23) - 127;
    const int f_exp_x = (int)((bx & 0x7F800000u) >> 23);
    const int newexp = f_exp_x - f_exp_y;
    const int satexp = newexp <= 0 ? 0 : newexp;
    a.m.head = __uint_as_float((by & 0x807FFFFFu) | 0x3F800000u);
    a.m.tail = __uint_as_float((bx & 0x807FFFFFu) | ((uint32_t)satexp << 23));
    a.e += f_exp_y;
}
FS_D Hdr<df32> reduced(Hdr<df32> a) { reduce(a); return a; }
// HDRFloat(U number), U = float / int  HDRFloat.h:295-363
FS_D Hdr<df32> hd_from_float(float x) {
    if (x == 0.0f) return hd_zero();
    const uint32_t b = __float_as_uint(x);
    return hd_make((int)((b & 0x7F800000u) >> 23) - 127, df_make(__uint_as_float((b & 0x807FFFFFu) | 0x3F800000u), 0.0f));
}
FS_D Hdr<df32> mul(Hdr<df32> a, Hdr<df32> b) { return hd_make(imax(a.e + b.e, MIN_BIG), df_mul(a.m, b.m)); }
FS_D Hdr<df32> square(Hdr<df32> a) { return hd_make(a.e * 2, df_mul(a.m, a.m)); } // square(): operator*, no clamp
// add_mutable / subtract_mutable  HDRFloat.h:974-1000, 1039-1065
template <bool Sub> FS_D Hdr<df32> hd_addsub(Hdr<df32> a, Hdr<df32> b) {
    const int d = a.e - b.e;
    if (d >= EXP_DIFF_IGNORED) return a;
    if (d >= 0) {
        const df32 t = df_mul(b.m, df_pow2neg(-d));
        a.m = Sub ? df_sub(a.m, t) : df_add(a.m, t);
    } else if (d > -EXP_DIFF_IGNORED) {
        const df32 t = df_mul(a.m, df_pow2neg(d));
        a.e = b.e;
        a.m = Sub ? df_sub(t, b.m) : df_add(t, b.m);
    } else {
        a.e = b.e;
        a.m = Sub ? df_neg(b.m) : b.m;
    }
    if (df_is_zero(a.m)) a.e = MIN_BIG;
    return a;
}
FS_D Hdr<df32> add(Hdr<df32> a, Hdr<df32> b) { return hd_addsub<false>(a, b); }
FS_D Hdr<df32> sub(Hdr<df32> a, Hdr<df32> b) { return hd_addsub<true>(a, b); }
// compareToBothPositiveReduced  HDRFloat.h:1150-1167
FS_D int cmp_pr(Hdr<df32> a, Hdr<df32> b) {
    if (a.e > b.e) return 1;
    if (a.e < b.e) return -1;
    if (df_gt(a.m, b.m)) return 1;
    if (df_lt(a.m, b.m)) return -1;
    return 0;
}
FS_D bool lt_pr(Hdr<df32> a, Hdr<df32> b) { return cmp_pr(a, b) < 0; }
FS_D bool ge_pr(Hdr<df32> a, Hdr<df32> b) { return cmp_pr(a, b) >= 0; }
FS_D bool gt_pr(Hdr<df32> a, Hdr<df32> b) { return cmp_pr(a, b) > 0; }
FS_D bool le_pr(Hdr<df32> a, Hdr<df32> b) { return cmp_pr(a, b) <= 0; }
// compareToBothPositiveReducedTemplate<256>() < 0  HDRFloat.h:1169-1184
FS_D bool lt_bailout(Hdr<df32> a) { return a.e < 1 || (a.e == 1 && !df_ge(a.m, df_from_float(256.0f))); }

// ---- HDRFloatComplex<CudaDblflt>  (HDRFloatComplex.h) ---------------------------------------------------------
FS_D HdrC<df32> hdc_zero() { HdrC<df32> r; r.re = df_make(0.0f, 0.0f); r.im = r.re; r.e = MIN_BIG; return r; }
// setMantexp  :158-173
FS_D HdrC<df32> hdc_from(Hdr<df32> re, Hdr<df32> im) {
    HdrC<df32> r;
    r.e = imax(re.e, im.e);
    r.re = df_mul(re.m, df_pow2(re.e - r.e));
    r.im = df_mul(im.m, df_pow2(im.e - r.e));
    return r;
}
// plus_mutable  :219-247
FS_D HdrC<df32> add(HdrC<df32> a, HdrC<df32> b) {
    const int d = a.e - b.e;
    if (d >= EXP_DIFF_IGNORED) return a;
    if (d >= 0) {
        const df32 m = df_pow2(-d);
        a.re = df_add(a.re, df_mul(b.re, m));
        a.im = df_add(a.im, df_mul(b.im, m));
    } else if (d > -EXP_DIFF_IGNORED) {
        const df32 m = df_pow2(d);
        a.e = b.e;
        a.re = df_add(df_mul(a.re, m), b.re);
        a.im = df_add(df_mul(a.im, m), b.im);
    } else {
        a = b;
    }
    return a;
}
// times_mutable  :267-283
FS_D HdrC<df32> mul(HdrC<df32> a, HdrC<df32> b) {
    HdrC<df32> r;
    r.re = df_sub(df_mul(a.re, b.re), df_mul(a.im, b.im));
    r.im = df_add(df_mul(a.re, b.im), df_mul(a.im, b.re));
    r.e = imax(a.e + b.e, MIN_BIG);
    return r;
}
// times_mutable(HDRFloat)  :334-348
FS_D HdrC<df32> mul(HdrC<df32> a, Hdr<df32> f) {
    HdrC<df32> r;
    r.re = df_mul(a.re, f.m);
    r.im = df_mul(a.im, f.m);
    r.e = imax(a.e + f.e, MIN_BIG);
    return r;
}
// Reduce  :473-527, 2x32 arm: two scalar Reduce + setMantexp
FS_D void reduce(HdrC<df32> &a) {
    if (df_is_zero(a.re) && df_is_zero(a.im)) return;
    // HDRFloat(CudaDblflt number): zero -> the zero value, else exponent 0 (HDRFloat.h:295-331)
    Hdr<df32> tr = df_is_zero(a.re) ? hd_zero() : hd_make(0, a.re);
    Hdr<df32> ti = df_is_zero(a.im) ? hd_zero() : hd_make(0, a.im);
    reduce(tr);
    reduce(ti);
    const int old = a.e;
    a = hdc_from(tr, ti);
    a.e += old;
}
FS_D Hdr<df32> cheb(HdrC<df32> a) {
    const Hdr<df32> r = hd_make(a.e, df_abs(a.re)), i = hd_make(a.e, df_abs(a.im));
    return cmp_pr(r, i) > 0 ? r : i;
}
FS_D Hdr<df32> norm2(HdrC<df32> a) { return hd_make(a.e << 1, df_add(df_mul(a.re, a.re), df_mul(a.im, a.im))); }

// ---- numeric policies -------------------------------------------------------------------------------------
struct Num2x32 {
    using Mant = df32;
    using Real = df32;
    using Cplx = Cx<df32>;
    static constexpr bool kHdr = false;
    static constexpr bool kDf = true;
    FS_D static Real zero() { return df_from_float(0.0f); }
    FS_D static Real from_int(int x) { return df_from_float((float)x); }
    FS_D static Real neg(Real a) { return df_neg(a); }
    FS_D static Cplx c_make(Real re, Real im) { Cplx c; c.re = re; c.im = im; return c; }
    FS_D static Cplx c_zero() { return c_make(zero(), zero()); }
    FS_D static Real c_re(Cplx c) { return c.re; }
    FS_D static Real c_im(Cplx c) { return c.im; }
    FS_D static Cplx c_mul2(Cplx c) { return mul(c, df_from_float(2.0f)); } // Ref * HDRFloat(2)
    FS_D static Real delta_x(Real dx, int X, Real centerX) { return df_sub(df_mul(dx, from_int(X)), centerX); }
    FS_D static Real delta_y(Real dy, int Y, Real centerY) { return df_sub(df_mul(df_neg(dy), from_int(Y)), centerY); }
    // LAKernel.cuh:141-176 (plain-type arm)
    FS_D static void perturb(Real &dx, Real &dy, Real zx, Real zy, Real cx, Real cy) {
        const Real two = df_from_float(2.0f);
        const Real mx2 = df_mul(zx, two), my2 = df_mul(zy, two);
        const Real s1 = df_add(my2, dy), s2 = df_add(mx2, dx);
        const Real nx = df_add(df_sub(df_mul(dx, s2), df_mul(dy, s1)), cx);
        const Real ny = df_add(df_add(df_mul(dx, s1), df_mul(dy, s2)), cy);
        dx = nx;
        dy = ny;
    }
    FS_D static Real norm2(Real x, Real y) { return df_add(df_mul(x, x), df_mul(y, y)); }
};

struct NumHdr2x32 {
    using Mant = df32;
    using Real = Hdr<df32>;
    using Cplx = HdrC<df32>;
    static constexpr bool kHdr = true;
    static constexpr bool kDf = true;
    FS_D static Real zero() { return hd_zero(); }
    FS_D static Real from_int(int x) { return hd_from_float((float)x); }
    FS_D static Real neg(Real a) { a.m = df_neg(a.m); return a; }
    FS_D static Cplx c_make(Real re, Real im) { return hdc_from(re, im); }
    FS_D static Cplx c_zero() { return hdc_from(hd_zero(), hd_zero()); } // TComplex{T(0), T(0)}
    FS_D static Real c_re(Cplx c) { return hd_make(c.e, c.re); }
    FS_D static Real c_im(Cplx c) { return hd_make(c.e, c.im); }
    FS_D static Cplx c_mul2(Cplx c) { return mul(c, hd_from_float(2.0f)); }
    FS_D static Real delta_x(Real dx, int X, Real centerX) { return sub(mul(dx, from_int(X)), centerX); }
    FS_D static Real delta_y(Real dy, int Y, Real centerY) { return sub(mul(neg(dy), from_int(Y)), centerY); }
    // custom_perturb3  HDRFloat.h:796-827
    FS_D static void perturb(Real &dx, Real &dy, Real zx, Real zy, Real cx, Real cy) {
        const Real two = hd_from_float(2.0f);
        const Real mx2 = mul(zx, two), my2 = mul(zy, two);
        const Real s1 = add(my2, dy), s2 = add(mx2, dx);
        Real nx = add(sub(mul(dx, s2), mul(dy, s1)), cx);
        reduce(nx);
        Real ny = add(add(mul(dx, s1), mul(dy, s2)), cy);
        reduce(ny);
        dx = nx;
        dy = ny;
    }
    FS_D static Real norm2(Real x, Real y) { return reduced(add(square(x), square(y))); }
};

} // namespace fs
