// fs_scaled_kernel.cuh -- "scaled" perturbation render kernel (row a5 of SURVEY.md section 8).
//
// Algorithm (what): FractalSharkGpuLib/ScaledKernels.cuh:3-239 `mandel_1x_float_perturb_scaled<IterType,T>`,
// launched by GPURenderer::RenderPerturbBLAScaled (GPU_Render.cu:1302-1377) for T = double
// (Gpu1x32PerturbedScaled) and T = HDRFloat<float> (GpuHDRx32PerturbedScaled).  The pixel delta is kept as
// w * S with w in plain binary32 and S in T; orbit elements flagged `bad` (tiny reference values) take one
// full step in T, everything else runs in binary32 and is re-scaled when |w|^2 grows past sqrt(1e30), on
// rebasing and at the end of the orbit.
//
// Rounding sequence: the binary32 step and the T = double arms follow the contraction nvcc chose for the
// reference build (sm_100a SASS of the kernel; identical for both instantiations):
//     2*X*x            -> p = x*X ; fma(x, X, p)                (NOT 2*p)
//     y*2 + twos*Y     -> fma(twos, Y, y + y)
//     s*X*X, s*Y*Y     -> fma(s*X, X, acc), fma(-Y, s*Y, acc)
//     x' + w*s         -> fma(s, w, x')        |.|^2 -> fma(a, a, b*b)        w2*s*s -> (s*w2)*s
// The T = HDRFloat<float> arms use the float+exponent operators of fs_types.cuh, one rounding per operator.
//
// Execution (how, B200-first): the persistent warp-tile queue shared with the other render kernels; the
// binary32 orbit keeps the reference's 16-byte `Bad` record {bad, pad, x, y}, which is exactly one LDG.128
// per step here (reference: three LDG.32), the T orbit its 24-byte record, read only on the rare T arms.
#pragma once
#include "fs_lav2.cuh"

namespace fs {

// GPUReferenceIter<float, PerturbExtras::Bad>  (GPU_ReferenceIter.h:10-21, 119-125): 16 bytes
struct alignas(16) ScaledElemF {
    uint32_t bad, pad;
    float x, y;
};
static_assert(sizeof(ScaledElemF) == 16, "GPUReferenceIter<float,Bad>");

template <class Num, class IterT> struct ScaledArgs {
    IterT *out;
    const ScaledElemF *orbit_f; // binary32 orbit
    const unsigned char *orbit_t; // GPUReferenceIter<T,Bad>[count], 24-byte records
    IterT orbit_count;
    int width, height, pitch;
    int shard_count, shard_index;
    typename Num::Real dx, dy, centerX, centerY;
    IterT n_iterations;
    TileQueue queue;
    unsigned long long *step_counter;
};

// ---- T-specific pieces ---------------------------------------------------------------------------------------
template <class Num> struct ScaledOps;

template <> struct ScaledOps<NumPlain<double>> {
    using T = double;
    FS_D static void load(const unsigned char *orbit, unsigned long long i, T &x, T &y) {
        const double *p = reinterpret_cast<const double *>(orbit + i * 24 + 8);
        x = __ldg(p);
        y = __ldg(p + 1);
    }
    FS_D static float to_float(T v) { return (float)v; }
    FS_D static T quot(T a, T b) { return a / b; }
    FS_D static T hypot_(T x, T y) { return sqrt(fma_(x, x, y * y)); } // HdrSqrt(x*x + y*y); HdrReduce no-op
    FS_D static T orbit_plus(T z, float w, T S) { return fma_(S, (double)w, z); } // z + (T)w * S
    FS_D static T times(float w, T S) { return S * (double)w; }                   // (T)w * S
    FS_D static T pixel_x(const T dx, int X, T cX) { return NumPlain<double>::delta_x(dx, X, cX); }
    FS_D static T pixel_y(const T dy, int Y, T cY) { return NumPlain<double>::delta_y(dy, Y, cY); }

    // one full step in T (ScaledKernels.cuh:159-233).  Returns false when the pixel escaped.
    template <class IterT>
    FS_D static bool full_step(const unsigned char *orbit, IterT &Ref, IterT last, T dR, T dI, T &S, float &X, float &Y,
                               T &newX, T &newY) {
        T xd, yd;
        load(orbit, Ref, xd, yd);
        const T Xo = (T)X, Yo = (T)Y;
        const T sX = S * Xo, sY = S * Yo;
        T a = fma_(Yo, yd, Yo * yd);       // Y*yd*2
        T b = fma_(Xo, xd, Xo * xd);       // X*xd*2
        b = b - a;
        b = fma_(Xo, sX, b);
        b = fma_(Yo, -sY, b);
        const T tx = b + dR / S;
        T c = fma_(Yo, S + S, yd + yd);    // yd*2 + 2*S*Y
        T d = fma_(Yo, xd, Yo * xd);       // Y*xd*2
        d = fma_(Xo, c, d);
        const T ty = d + dI / S;
        ++Ref;
        load(orbit, Ref, xd, yd);
        const T px = S * tx, py = S * ty;
        const T zy = yd + py, zx = xd + px;
        const T zn = fma_(zx, zx, zy * zy);
        if (!(zn < 256.0)) return false;
        const T SS = S * S;
        const T nrm = fma_(SS, tx * tx, SS * (ty * ty));
        if (zn < nrm || Ref == last) {
            newX = zx; newY = zy; Ref = 0;
        } else {
            newX = px; newY = py;
        }
        return true;
    }
};

template <> struct ScaledOps<NumHdr<float>> {
    using T = Hdr<float>;
    // x is stored Left-order {m, e}, y Right-order {e, m} (GPU_ReferenceIter.h:119-125)
    FS_D static void load(const unsigned char *orbit, unsigned long long i, T &x, T &y) {
        const uint2 *p = reinterpret_cast<const uint2 *>(orbit + i * 24 + 8);
        const uint2 a = __ldg(p), b = __ldg(p + 1);
        x.m = __uint_as_float(a.x); x.e = (int)a.y;
        y.e = (int)b.x; y.m = __uint_as_float(b.y);
    }
    FS_D static float to_float(T v) { return v.m * MT<float>::pow2(v.e); } // toDouble()  HDRFloat.h:553-557
    FS_D static T quot(T a, T b) { return div(a, b); }
    // HdrSqrt  HDRFloat.h:1377-1382 (exponent halved, odd exponents fold a factor 2 into the mantissa)
    FS_D static T hdr_sqrt(T v) {
        const bool odd = (v.e & 1) != 0;
        T r;
        r.e = odd ? (v.e - 1) / 2 : v.e / 2;
        r.m = sqrtf(odd ? 2.0f * v.m : v.m);
        return r;
    }
    FS_D static T hypot_(T x, T y) { return reduced(hdr_sqrt(add(mul(x, x), mul(y, y)))); }
    FS_D static T orbit_plus(T z, float w, T S) { return add(z, mul(hdr_from<float>(w), S)); }
    FS_D static T times(float w, T S) { return mul(hdr_from<float>(w), S); }
    FS_D static T pixel_x(const T dx, int X, T cX) { return reduced(NumHdr<float>::delta_x(dx, X, cX)); }
    FS_D static T pixel_y(const T dy, int Y, T cY) { return reduced(NumHdr<float>::delta_y(dy, Y, cY)); }

    template <class IterT>
    FS_D static bool full_step(const unsigned char *orbit, IterT &Ref, IterT last, T dR, T dI, T &S, float &X, float &Y,
                               T &newX, T &newY) {
        T xd, yd;
        load(orbit, Ref, xd, yd);
        const T Xo = hdr_from<float>(X), Yo = hdr_from<float>(Y);
        T tx = mul2(mul(Xo, xd));
        tx = sub(tx, mul2(mul(Yo, yd)));
        tx = add(tx, mul(mul(S, Xo), Xo));
        tx = sub(tx, mul(mul(S, Yo), Yo));
        tx = add(tx, div(dR, S));
        reduce(tx);
        T ty = mul(Xo, add(mul2(yd), mul(mul(hdr_make<float>(1, 1.0f), S), Yo)));
        ty = add(ty, mul2(mul(Yo, xd)));
        ty = add(ty, div(dI, S));
        reduce(ty);
        ++Ref;
        load(orbit, Ref, xd, yd);
        const T zx = add(xd, mul(tx, S)), zy = add(yd, mul(ty, S));
        const T zn = reduced(add(mul(zx, zx), mul(zy, zy)));
        if (!lt_bailout(zn)) return false;
        const T SS = mul(S, S);
        const T nrm = reduced(add(mul(mul(tx, tx), SS), mul(mul(ty, ty), SS)));
        if (lt_pr(zn, nrm) || Ref == last) {
            newX = add(xd, mul(tx, S)); newY = add(yd, mul(ty, S)); Ref = 0;
        } else {
            newX = mul(tx, S); newY = mul(ty, S);
        }
        return true;
    }
};

// ---- one pixel -------------------------------------------------------------------------------------------------
template <class Num, class IterT, bool Count>
FS_D IterT scaled_pixel(const ScaledArgs<Num, IterT> &A, int Xp, int Yp, unsigned long long &steps) {
    using Ops = ScaledOps<Num>;
    using T = typename Num::Real;
    const float LARGE_MANTISSA = 1e30;
    const float w2threshold = exp(log(LARGE_MANTISSA) / 2);

    IterT iter = 0, Ref = 0;
    const T dR = Ops::pixel_x(A.dx, Xp, A.centerX);
    const T dI = Ops::pixel_y(A.dy, Yp, A.centerY);
    T S = Ops::hypot_(dR, dI);
    float c0x = Ops::to_float(Ops::quot(dR, S));
    float c0y = Ops::to_float(Ops::quot(dI, S));
    float X = 0, Y = 0;
    float s = Ops::to_float(S);
    float twos = s + s;
    const IterT last = A.orbit_count - 1;

    auto rescale = [&](T nx, T ny) {
        S = Ops::hypot_(nx, ny);
        s = Ops::to_float(S);
        twos = s + s;
        c0x = Ops::to_float(Ops::quot(dR, S));
        c0y = Ops::to_float(Ops::quot(dI, S));
        X = Ops::to_float(Ops::quot(nx, S));
        Y = Ops::to_float(Ops::quot(ny, S));
    };

    // Eight binary32 steps per branch where nothing can happen: every step's tests (escape, rebase by norm, |w|^2
    // past the re-scaling threshold, a `bad` element; the end of the orbit is excluded by the chunk's range) are still
    // evaluated and folded into one predicate; a chunk in which any of them fired is discarded and the steps are
    // taken one by one by the loop below, which is the reference's loop.  One LDG.128 per step (the element a step
    // arrives at is the one the next step starts from) and ~30 instead of ~40 instructions per step.
    constexpr int K = 8;
    int exact_budget = 0; // steps to take one by one after a discarded chunk (whatever fired is within K steps)
    while (iter < A.n_iterations) {
        // A chunk is attempted only where it can plausibly complete: not within K steps of the end of the orbit or
        // of the iteration limit, not right after a discarded chunk, and not while |w|^2 is within a factor 2^20 of
        // the re-scaling threshold (|w| rarely grows by more than ~2.4x per step, so eight steps rarely bridge that gap; pixels
        // whose w races to the threshold every couple of dozen steps would otherwise throw every other chunk away).
        // These are performance choices only: what a chunk computes is checked step by step either way.
        const bool can = exact_budget == 0 && (uint64_t)iter + K <= (uint64_t)A.n_iterations &&
                         (uint64_t)Ref + K < (uint64_t)last && fma_(X, X, Y * Y) < 0x1p30f;
        if (can) {
            const float X0 = X, Y0 = Y;
            ScaledElemF e = ldg_rec(A.orbit_f + Ref);
            bool ok = true;
#pragma unroll
            for (int u = 0; u < K; u++) {
                ok = ok && (e.bad == 0);
                const float sX = s * X, sY = s * Y;
                float a = fma_(X, e.x, X * e.x);
                const float b = fma_(Y, e.y, Y * e.y);
                const float yx = fma_(Y, e.x, Y * e.x);
                const float t = fma_(twos, Y, e.y + e.y);
                a = a - b;
                const float ny = fma_(X, t, yx);
                a = fma_(X, sX, a);
                a = fma_(-Y, sY, a);
                Y = c0y + ny;
                X = c0x + a;
                e = ldg_rec(A.orbit_f + Ref + (u + 1));
                const float zy = fma_(s, Y, e.y), zx = fma_(s, X, e.x);
                const float w2 = fma_(X, X, Y * Y);
                const float zn = fma_(zx, zx, zy * zy);
                const float nrm = s * (s * w2);
                ok = ok && (zn < 256.0f) && !(zn < nrm) && !(w2 >= w2threshold);
            }
            if (ok) {
                Ref += (IterT)K;
                iter += (IterT)K;
                if (Count) steps += K;
            } else {
                X = X0;
                Y = Y0;
                exact_budget = K;
            }
            continue;
        }
        if (exact_budget > 0) exact_budget--;
        const ScaledElemF e = ldg_rec(A.orbit_f + Ref);
        if (Count) steps++;
        if (e.bad == 0) {
            const float sX = s * X, sY = s * Y;
            float a = fma_(X, e.x, X * e.x);   // X*x*2
            const float b = fma_(Y, e.y, Y * e.y); // Y*y*2
            const float yx = fma_(Y, e.x, Y * e.x); // Y*x*2
            const float t = fma_(twos, Y, e.y + e.y);
            a = a - b;
            const float ny = fma_(X, t, yx);
            a = fma_(X, sX, a);
            a = fma_(-Y, sY, a);
            Y = c0y + ny;
            X = c0x + a;
            ++Ref;
            const ScaledElemF n = ldg_rec(A.orbit_f + Ref);
            const float zy = fma_(s, Y, n.y), zx = fma_(s, X, n.x);
            const float w2 = fma_(X, X, Y * Y);
            const float zn = fma_(zx, zx, zy * zy);
            const float nrm = s * (s * w2);
            const bool zn_ok = zn < 256.0f;
            const bool test1ab = (zn < nrm) || (Ref == last && zn_ok);
            const bool testw2 = (w2 >= w2threshold) && zn_ok;
            if (!test1ab && !testw2 && zn_ok) {
                ++iter;
                continue;
            } else if (test1ab) {
                T xd, yd;
                Ops::load(A.orbit_t, Ref, xd, yd);
                const T nx = Ops::orbit_plus(xd, X, S), nyT = Ops::orbit_plus(yd, Y, S);
                Ref = 0;
                rescale(nx, nyT);
                ++iter;
                continue;
            } else if (testw2) {
                const T nx = Ops::times(X, S), nyT = Ops::times(Y, S);
                rescale(nx, nyT);
                ++iter;
                continue;
            } else {
                break;
            }
        } else {
            T nx, nyT;
            if (!Ops::template full_step<IterT>(A.orbit_t, Ref, last, dR, dI, S, X, Y, nx, nyT)) break;
            rescale(nx, nyT);
        }
        ++iter;
    }
    return iter;
}

template <class Num, class IterT, bool Count>
__global__ void __launch_bounds__(256) scaled_kernel(const ScaledArgs<Num, IterT> A) {
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
        if (X < A.width && Y < A.height) A.out[(size_t)Y * A.pitch + X] = scaled_pixel<Num, IterT, Count>(A, X, Y, steps);
        __syncwarp();
    }
    if (Count && A.step_counter) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(A.step_counter, steps);
    }
}

} // namespace fs
