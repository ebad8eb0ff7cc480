// fs_orbit_rc.cuh -- compressed reference orbits (PerturbExtras::SimpleCompression, the `RC` algorithms; the
// compressed half of row a2 of SURVEY.md section 8).
//
// What: FractalSharkGpuLib/Perturb.cuh:160-206 (SeqWorkspace), :246-326 (BinarySearch, GetCompressedComplex,
// GetCompressedComplexSeq).  A compressed orbit is a list of waypoints {CompressionIndex, x, y}; the element at
// uncompressed index i is the nearest waypoint at or below i, advanced i - CompressionIndex times by
//     zx' = zx*zx - zy*zy + OrbitXLow ; zy' = T{2}*zx_old*zy + OrbitYLow     (each HdrReduce'd)
// in the low type T.  That value depends on nothing but the waypoint, so every access pattern of the reference
// (sequential walk, restart after a rebase, binary search + replay) sees the same element values.
//
// How (B200-first): the reference replays on the fly in every thread of every pixel.  With 180 GB of HBM the
// uncompressed orbit of even a 2^31-iteration reference fits (16 B x 2^31 = 34 GB), so the waypoint list is
// expanded ONCE per upload by `orbit_expand_kernel` -- one thread per waypoint segment, all segments in parallel
// -- into the ordinary uncompressed layout, and the render kernels (including the HDRx32 scaled-chunk path and
// its per-element table) run unchanged.  Host->device traffic stays the compressed size.
//
// Rounding of the replay step: float+exponent and 2x32 types evaluate operator by operator (one rounding each);
// plain float/double follow the contraction in the reference's sm_100a SASS of the RC kernels:
//     t = zy*zy ; zx' = fma(zx, zx, -t) + X ; zy' = fma(zx + zx, zy, Y).
#pragma once
#include "fs_lav2.cuh"

namespace fs {

template <class Num> struct RcStep;
template <class M> struct RcStep<NumPlain<M>> {
    FS_D static void run(M &zx, M &zy, M X, M Y) {
        const M t = zy * zy;
        const M two_x = zx + zx;
        const M nx = fma_(zx, zx, -t) + X;
        zy = fma_(two_x, zy, Y);
        zx = nx;
    }
};
template <class M> struct RcStep<NumHdr<M>> {
    FS_D static void run(Hdr<M> &zx, Hdr<M> &zy, Hdr<M> X, Hdr<M> Y) {
        const Hdr<M> o = zx;
        zx = reduced(add(sub(mul(zx, zx), mul(zy, zy)), X));
        zy = reduced(add(mul(mul(hdr_make<M>(1, M(1)), o), zy), Y));
    }
};
template <> struct RcStep<Num2x32> {
    FS_D static void run(df32 &zx, df32 &zy, df32 X, df32 Y) {
        const df32 o = zx;
        zx = df_add(df_sub(df_mul(zx, zx), df_mul(zy, zy)), X);
        zy = df_add(df_mul(df_mul(df_from_float(2.0f), o), zy), Y);
    }
};
template <> struct RcStep<NumHdr2x32> {
    FS_D static void run(Hdr<df32> &zx, Hdr<df32> &zy, Hdr<df32> X, Hdr<df32> Y) {
        const Hdr<df32> o = zx;
        zx = reduced(add(sub(mul(zx, zx), mul(zy, zy)), X));
        zy = reduced(add(mul(mul(hd_from_float(2.0f), o), zy), Y));
    }
};

// element body of a waypoint record (8-byte aligned, same field order as the uncompressed record)
template <class Num> struct RcLoad;
template <> struct RcLoad<NumPlain<float>> {
    FS_D static void get(const unsigned char *p, float &x, float &y) {
        const float *f = reinterpret_cast<const float *>(p);
        x = f[0]; y = f[1];
    }
};
template <> struct RcLoad<NumPlain<double>> {
    FS_D static void get(const unsigned char *p, double &x, double &y) {
        const double *d = reinterpret_cast<const double *>(p);
        x = d[0]; y = d[1];
    }
};
template <> struct RcLoad<NumHdr<float>> {
    FS_D static void get(const unsigned char *p, Hdr<float> &x, Hdr<float> &y) {
        const uint32_t *u = reinterpret_cast<const uint32_t *>(p);
        x.m = __uint_as_float(u[0]); x.e = (int)u[1]; y.e = (int)u[2]; y.m = __uint_as_float(u[3]);
    }
};
template <> struct RcLoad<NumHdr<double>> {
    FS_D static void get(const unsigned char *p, Hdr<double> &x, Hdr<double> &y) {
        x.m = *reinterpret_cast<const double *>(p); x.e = *reinterpret_cast<const int *>(p + 8);
        y.e = *reinterpret_cast<const int *>(p + 16); y.m = *reinterpret_cast<const double *>(p + 24);
    }
};
template <> struct RcLoad<Num2x32> {
    FS_D static void get(const unsigned char *p, df32 &x, df32 &y) {
        const float *f = reinterpret_cast<const float *>(p);
        x.head = f[0]; x.tail = f[1]; y.head = f[2]; y.tail = f[3];
    }
};
template <> struct RcLoad<NumHdr2x32> {
    FS_D static void get(const unsigned char *p, Hdr<df32> &x, Hdr<df32> &y) {
        const uint32_t *u = reinterpret_cast<const uint32_t *>(p);
        x.m.head = __uint_as_float(u[0]); x.m.tail = __uint_as_float(u[1]); x.e = (int)u[2];
        y.e = (int)u[3]; y.m.head = __uint_as_float(u[4]); y.m.tail = __uint_as_float(u[5]);
    }
};

// element store in the uncompressed layouts OrbitIO<Num>::load reads
template <class Num> struct OrbitStore;
template <> struct OrbitStore<NumPlain<float>> {
    FS_D static void put(void *base, uint64_t i, float x, float y) { reinterpret_cast<float2 *>(base)[i] = make_float2(x, y); }
};
template <> struct OrbitStore<NumPlain<double>> {
    FS_D static void put(void *base, uint64_t i, double x, double y) { reinterpret_cast<double2 *>(base)[i] = make_double2(x, y); }
};
template <> struct OrbitStore<NumHdr<float>> {
    FS_D static void put(void *base, uint64_t i, Hdr<float> x, Hdr<float> y) {
        reinterpret_cast<uint4 *>(base)[i] = make_uint4(__float_as_uint(x.m), (unsigned)x.e, (unsigned)y.e, __float_as_uint(y.m));
    }
};
template <> struct OrbitStore<NumHdr<double>> {
    FS_D static void put(void *base, uint64_t i, Hdr<double> x, Hdr<double> y) {
        uint4 *p = reinterpret_cast<uint4 *>(base) + 2 * i;
        p[0] = make_uint4((unsigned)__double2loint(x.m), (unsigned)__double2hiint(x.m), (unsigned)x.e, 0u);
        p[1] = make_uint4((unsigned)y.e, 0u, (unsigned)__double2loint(y.m), (unsigned)__double2hiint(y.m));
    }
};
template <> struct OrbitStore<Num2x32> {
    FS_D static void put(void *base, uint64_t i, df32 x, df32 y) {
        reinterpret_cast<float4 *>(base)[i] = make_float4(x.head, x.tail, y.head, y.tail);
    }
};
template <> struct OrbitStore<NumHdr2x32> {
    FS_D static void put(void *base, uint64_t i, Hdr<df32> x, Hdr<df32> y) {
        uint2 *p = reinterpret_cast<uint2 *>(base) + 3 * i;
        p[0] = make_uint2(__float_as_uint(x.m.head), __float_as_uint(x.m.tail));
        p[1] = make_uint2((unsigned)x.e, (unsigned)y.e);
        p[2] = make_uint2(__float_as_uint(y.m.head), __float_as_uint(y.m.tail));
    }
};

// One thread per waypoint: writes the waypoint, then replays up to (not including) the next waypoint's index.
// `wire` = GPUReferenceIter<T, SimpleCompression>[n_way]: 8-byte {CompressionIndex : 63, Rebase : 1} prefix followed
// by the element in the same byte layout as the uncompressed record (GPU_ReferenceIter.h:24-49, 119-125).
template <class Num>
__global__ void __launch_bounds__(128) orbit_expand_kernel(const unsigned char *__restrict__ wire, uint64_t n_way,
                                                           uint64_t n_full, typename Num::Real X, typename Num::Real Y,
                                                           void *__restrict__ out) {
    using Real = typename Num::Real;
    constexpr size_t kStride = 8 + OrbitIO<Num>::kBytes;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_way; k += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned char *rec = wire + k * kStride;
        const uint64_t begin = *reinterpret_cast<const uint64_t *>(rec) & 0x7FFFFFFFFFFFFFFFull;
        uint64_t end = n_full;
        if (k + 1 < n_way) end = *reinterpret_cast<const uint64_t *>(rec + kStride) & 0x7FFFFFFFFFFFFFFFull;
        if (end > n_full) end = n_full;
        Real zx, zy;
        RcLoad<Num>::get(rec + 8, zx, zy);
        for (uint64_t i = begin; i < end; i++) {
            OrbitStore<Num>::put(out, i, zx, zy);
            RcStep<Num>::run(zx, zy, X, Y);
        }
    }
}

} // namespace fs
