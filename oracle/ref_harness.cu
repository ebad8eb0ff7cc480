// ref_harness.cu -- TEST INFRASTRUCTURE ONLY (Oracle A of SURVEY.md section 8c).
//
// A flat C interface around the UNMODIFIED reference `GPURenderer` (FractalSharkLib/GPU_Render.h,
// object code built from /root/reference/FractalSharkGpuLib/GPU_Render.cu for sm_100a by
// oracle/Makefile into oracle/_ref/libref_gpurender.so).  It lets tests/ and bench.py's
// `--impl reference` arm run the reference's own kernels on the same B200 with the same inputs
// as libfsgpu.so and compare iteration buffers element-wise.
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference arm may load this library; the
// product (fractalshark_b200/) never does.  No reference source is copied: this file only
// #includes the reference headers where they lie and calls their public entry points.  Host-side
// `LAReference` objects are synthesised from the flat tables by writing their (private) fields --
// the reference builds them inside RefOrbitCalc, which is outside the hot path.
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#define private public
#define protected public
#include "GPU_Render.h"
#include "dblflt.cuh"
#include "CudaDblflt.h"
#include "HDRFloat.h"
#include "HDRFloatComplex.h"
#include "GPU_LAReference.h"
#include "GPU_LAInfoDeep.h"
#include "LAReference.h"
#undef private
#undef protected

#include <cuda_runtime.h>

// GPU_Render.o leaves exactly these GrowableVector members undefined (nm -C --undefined-only);
// the reference defines them in HpSharkFloatLib/Vectors.cpp, which drags in the file-mapping layer.
template <class EltT> EltT *GrowableVector<EltT>::GetData() const { return m_Data; }
template <class EltT> size_t GrowableVector<EltT>::GetSize() const { return m_UsedSizeInElts; }

#define REFH_INST(EltT)                                                                                                \
    template EltT *GrowableVector<EltT>::GetData() const;                                                              \
    template size_t GrowableVector<EltT>::GetSize() const;
REFH_INST(LAStageInfo<uint32_t>)
REFH_INST(LAStageInfo<uint64_t>)
#define REFH_ALL_T(IterType)                                                                                           \
    template LAInfoDeep<IterType, float, float, PerturbExtras::Disable> *GrowableVector<LAInfoDeep<IterType, float, float, PerturbExtras::Disable>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, float, float, PerturbExtras::Disable>>::GetSize() const;       \
    template LAInfoDeep<IterType, float, float, PerturbExtras::SimpleCompression> *GrowableVector<LAInfoDeep<IterType, float, float, PerturbExtras::SimpleCompression>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, float, float, PerturbExtras::SimpleCompression>>::GetSize() const; \
    template LAInfoDeep<IterType, double, double, PerturbExtras::Disable> *GrowableVector<LAInfoDeep<IterType, double, double, PerturbExtras::Disable>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, double, double, PerturbExtras::Disable>>::GetSize() const;     \
    template LAInfoDeep<IterType, double, double, PerturbExtras::SimpleCompression> *GrowableVector<LAInfoDeep<IterType, double, double, PerturbExtras::SimpleCompression>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, double, double, PerturbExtras::SimpleCompression>>::GetSize() const; \
    template LAInfoDeep<IterType, CudaDblflt<MattDblflt>, CudaDblflt<MattDblflt>, PerturbExtras::Disable> *GrowableVector<LAInfoDeep<IterType, CudaDblflt<MattDblflt>, CudaDblflt<MattDblflt>, PerturbExtras::Disable>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, CudaDblflt<MattDblflt>, CudaDblflt<MattDblflt>, PerturbExtras::Disable>>::GetSize() const; \
    template LAInfoDeep<IterType, CudaDblflt<MattDblflt>, CudaDblflt<MattDblflt>, PerturbExtras::SimpleCompression> *GrowableVector<LAInfoDeep<IterType, CudaDblflt<MattDblflt>, CudaDblflt<MattDblflt>, PerturbExtras::SimpleCompression>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, CudaDblflt<MattDblflt>, CudaDblflt<MattDblflt>, PerturbExtras::SimpleCompression>>::GetSize() const; \
    template LAInfoDeep<IterType, HDRFloat<float>, float, PerturbExtras::Disable> *GrowableVector<LAInfoDeep<IterType, HDRFloat<float>, float, PerturbExtras::Disable>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, HDRFloat<float>, float, PerturbExtras::Disable>>::GetSize() const; \
    template LAInfoDeep<IterType, HDRFloat<float>, float, PerturbExtras::SimpleCompression> *GrowableVector<LAInfoDeep<IterType, HDRFloat<float>, float, PerturbExtras::SimpleCompression>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, HDRFloat<float>, float, PerturbExtras::SimpleCompression>>::GetSize() const; \
    template LAInfoDeep<IterType, HDRFloat<double>, double, PerturbExtras::Disable> *GrowableVector<LAInfoDeep<IterType, HDRFloat<double>, double, PerturbExtras::Disable>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, HDRFloat<double>, double, PerturbExtras::Disable>>::GetSize() const; \
    template LAInfoDeep<IterType, HDRFloat<double>, double, PerturbExtras::SimpleCompression> *GrowableVector<LAInfoDeep<IterType, HDRFloat<double>, double, PerturbExtras::SimpleCompression>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, HDRFloat<double>, double, PerturbExtras::SimpleCompression>>::GetSize() const; \
    template LAInfoDeep<IterType, HDRFloat<CudaDblflt<MattDblflt>>, CudaDblflt<MattDblflt>, PerturbExtras::Disable> *GrowableVector<LAInfoDeep<IterType, HDRFloat<CudaDblflt<MattDblflt>>, CudaDblflt<MattDblflt>, PerturbExtras::Disable>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, HDRFloat<CudaDblflt<MattDblflt>>, CudaDblflt<MattDblflt>, PerturbExtras::Disable>>::GetSize() const; \
    template LAInfoDeep<IterType, HDRFloat<CudaDblflt<MattDblflt>>, CudaDblflt<MattDblflt>, PerturbExtras::SimpleCompression> *GrowableVector<LAInfoDeep<IterType, HDRFloat<CudaDblflt<MattDblflt>>, CudaDblflt<MattDblflt>, PerturbExtras::SimpleCompression>>::GetData() const; \
    template size_t GrowableVector<LAInfoDeep<IterType, HDRFloat<CudaDblflt<MattDblflt>>, CudaDblflt<MattDblflt>, PerturbExtras::SimpleCompression>>::GetSize() const;
REFH_ALL_T(uint32_t)
REFH_ALL_T(uint64_t)

namespace {

struct Harness {
    GPURenderer renderer;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<unsigned char> la_storage; // backing store of the synthesised LAReference object
};

// Build a host LAReference<...> image without running its constructor: only the members read by
// GPU_LAReference's upload constructor (GPU_LAReference.h:78-162) are populated.
template <typename IterType, class T, class SubType, PerturbExtras PExtras = PerturbExtras::Disable>
const LAReference<IterType, T, SubType, PExtras> *
make_la(Harness *h, const void *las, uint64_t num_las, const void *stages, uint64_t num_stages, const void *at,
        uint64_t stage_count, int use_at, int is_valid) {
    using LR = LAReference<IterType, T, SubType, PExtras>;
    h->la_storage.assign(sizeof(LR) + 64, 0);
    unsigned char *p = h->la_storage.data();
    p += (64 - (reinterpret_cast<uintptr_t>(p) & 63)) & 63;
    LR *lr = reinterpret_cast<LR *>(p);
    lr->m_UseAT = use_at != 0;
    if (at) memcpy(&lr->m_AT, at, sizeof(lr->m_AT));
    lr->m_LAStageCount = (IterType)stage_count;
    lr->m_IsValid = is_valid != 0;
    lr->m_LAs.m_Data = (LAInfoDeep<IterType, T, SubType, PExtras> *)las;
    lr->m_LAs.m_UsedSizeInElts = num_las;
    lr->m_LAs.m_CapacityInElts = num_las;
    lr->m_LAStages.m_Data = (LAStageInfo<IterType> *)stages;
    lr->m_LAStages.m_UsedSizeInElts = num_stages;
    lr->m_LAStages.m_CapacityInElts = num_stages;
    return lr;
}

template <class T> T pod(const void *p) {
    T v{};
    memcpy((void *)&v, p, sizeof(T));
    return v;
}

template <typename IterType, class T, class SubType>
uint32_t init_perturb_t(Harness *h, uint64_t gen, const void *orbit, uint64_t count, uint64_t period, const void *xlow,
                        const void *ylow, const void *las, uint64_t num_las, const void *stages, uint64_t num_stages,
                        const void *at, uint64_t stage_count, int use_at, int is_valid) {
    GPUPerturbResults<IterType, T, PerturbExtras::Disable> res{
        (IterType)count, (IterType)count, xlow ? pod<T>(xlow) : T{}, ylow ? pod<T>(ylow) : T{},
        (const GPUReferenceIter<T, PerturbExtras::Disable> *)orbit, (IterType)period};
    const LAReference<IterType, T, SubType, PerturbExtras::Disable> *la = nullptr;
    if (las) la = make_la<IterType, T, SubType>(h, las, num_las, stages, num_stages, at, stage_count, use_at, is_valid);
    return h->renderer.InitializePerturb<IterType, T, SubType, PerturbExtras::Disable, T>(gen, &res, 0, nullptr, la);
}

// compressed orbit (PerturbExtras::SimpleCompression): `count` waypoints standing for `full` orbit entries
template <typename IterType, class T, class SubType>
uint32_t init_perturb_rc_t(Harness *h, uint64_t gen, const void *orbit, uint64_t count, uint64_t full, uint64_t period,
                           const void *xlow, const void *ylow, const void *las, uint64_t num_las, const void *stages,
                           uint64_t num_stages, const void *at, uint64_t stage_count, int use_at, int is_valid) {
    constexpr PerturbExtras PE = PerturbExtras::SimpleCompression;
    GPUPerturbResults<IterType, T, PE> res{(IterType)count, (IterType)full, pod<T>(xlow), pod<T>(ylow),
                                           (const GPUReferenceIter<T, PE> *)orbit, (IterType)period};
    const LAReference<IterType, T, SubType, PE> *la = nullptr;
    if (las) la = make_la<IterType, T, SubType, PE>(h, las, num_las, stages, num_stages, at, stage_count, use_at, is_valid);
    return h->renderer.InitializePerturb<IterType, T, SubType, PE, T>(gen, &res, 0, nullptr, la);
}

template <typename IterType, class T, class SubType>
uint32_t render_lav2_rc_t(Harness *h, uint32_t alg, int mode, const void *cx, const void *cy, const void *dx,
                          const void *dy, const void *cenx, const void *ceny, uint64_t n) {
    constexpr PerturbExtras PE = PerturbExtras::SimpleCompression;
    RenderAlgorithm a; *const_cast<RenderAlgorithmEnum *>(&a.Algorithm) = (RenderAlgorithmEnum)alg;
    const T vcx = pod<T>(cx), vcy = pod<T>(cy), vdx = pod<T>(dx), vdy = pod<T>(dy), vx = pod<T>(cenx), vy = pod<T>(ceny);
    switch (mode) {
    case 1: return h->renderer.RenderPerturbLAv2<IterType, T, SubType, LAv2Mode::Full, PE>(a, vcx, vcy, vdx, vdy, vx, vy, (IterType)n);
    case 2: return h->renderer.RenderPerturbLAv2<IterType, T, SubType, LAv2Mode::PO, PE>(a, vcx, vcy, vdx, vdy, vx, vy, (IterType)n);
    case 3: return h->renderer.RenderPerturbLAv2<IterType, T, SubType, LAv2Mode::LAO, PE>(a, vcx, vcy, vdx, vdy, vx, vy, (IterType)n);
    default: return 10100;
    }
}

template <typename IterType, class T, class SubType>
uint32_t render_lav2_t(Harness *h, uint32_t alg, int mode, const void *cx, const void *cy, const void *dx,
                       const void *dy, const void *cenx, const void *ceny, uint64_t n) {
    RenderAlgorithm a; *const_cast<RenderAlgorithmEnum *>(&a.Algorithm) = (RenderAlgorithmEnum)alg;
    const T vcx = pod<T>(cx), vcy = pod<T>(cy), vdx = pod<T>(dx), vdy = pod<T>(dy), vx = pod<T>(cenx), vy = pod<T>(ceny);
    switch (mode) {
    case 1: return h->renderer.RenderPerturbLAv2<IterType, T, SubType, LAv2Mode::Full, PerturbExtras::Disable>(a, vcx, vcy, vdx, vdy, vx, vy, (IterType)n);
    case 2: return h->renderer.RenderPerturbLAv2<IterType, T, SubType, LAv2Mode::PO, PerturbExtras::Disable>(a, vcx, vcy, vdx, vdy, vx, vy, (IterType)n);
    case 3: return h->renderer.RenderPerturbLAv2<IterType, T, SubType, LAv2Mode::LAO, PerturbExtras::Disable>(a, vcx, vcy, vdx, vdy, vx, vy, (IterType)n);
    default: return 10100;
    }
}

// RenderPerturbBLA reads only BLAS::m_B and BLAS::m_LM2 (GPU_Render.cu:1478-1483, BLA.cuh:288-386): a BLAS object is
// synthesised from the flat per-level tables without running its constructor (which needs PerturbationResults).
template <typename IterType, class T>
uint32_t render_bla_t(Harness *h, uint32_t alg, const void *orbit, uint64_t count, uint64_t period,
                      const void *const *levels, const uint64_t *counts, uint32_t num_levels, int lm2, const void *cx,
                      const void *cy, const void *dx, const void *dy, const void *cenx, const void *ceny, uint64_t n) {
    RenderAlgorithm a; *const_cast<RenderAlgorithmEnum *>(&a.Algorithm) = (RenderAlgorithmEnum)alg;
    GPUPerturbResults<IterType, T, PerturbExtras::Disable> res{
        (IterType)count, (IterType)count, T{}, T{}, (const GPUReferenceIter<T, PerturbExtras::Disable> *)orbit, (IterType)period};
    using BL = BLAS<IterType, T>;
    std::vector<unsigned char> storage(sizeof(BL) + 64, 0);
    unsigned char *p = storage.data();
    p += (64 - (reinterpret_cast<uintptr_t>(p) & 63)) & 63;
    BL *b = reinterpret_cast<BL *>(p);
    new (&b->m_B) std::vector<std::vector<BLA<T>>>(num_levels);
    for (uint32_t i = 0; i < num_levels; i++) {
        if (!levels[i] || !counts[i]) continue;
        b->m_B[i].resize(counts[i]);
        memcpy((void *)b->m_B[i].data(), levels[i], counts[i] * sizeof(BLA<T>));
    }
    b->m_LM2 = lm2;
    const T vcx = pod<T>(cx), vcy = pod<T>(cy), vdx = pod<T>(dx), vdy = pod<T>(dy), vx = pod<T>(cenx), vy = pod<T>(ceny);
    cudaEventRecord(h->ev0, h->renderer.m_ComputeStream);
    const uint32_t rc = h->renderer.RenderPerturbBLA<IterType, T>(a, &res, b, vcx, vcy, vdx, vdy, vx, vy, (IterType)n, 1);
    cudaEventRecord(h->ev1, h->renderer.m_ComputeStream);
    h->renderer.SyncComputeStream();
    using VV = std::vector<std::vector<BLA<T>>>;
    b->m_B.~VV();
    return rc;
}

// GPURenderer::RenderPerturbBLAScaled<IterType,T>: both orbits in the Bad layout, uploaded per call
template <typename IterType, class T>
uint32_t render_scaled_t(Harness *h, uint32_t alg, const void *orbit_t, const void *orbit_f, uint64_t count,
                         uint64_t period, const void *cx, const void *cy, const void *dx, const void *dy,
                         const void *cenx, const void *ceny, uint64_t n) {
    RenderAlgorithm a; *const_cast<RenderAlgorithmEnum *>(&a.Algorithm) = (RenderAlgorithmEnum)alg;
    GPUPerturbResults<IterType, T, PerturbExtras::Bad> rt{
        (IterType)count, (IterType)count, T{}, T{}, (const GPUReferenceIter<T, PerturbExtras::Bad> *)orbit_t, (IterType)period};
    GPUPerturbResults<IterType, float, PerturbExtras::Bad> rf{
        (IterType)count, (IterType)count, 0.0f, 0.0f, (const GPUReferenceIter<float, PerturbExtras::Bad> *)orbit_f, (IterType)period};
    const T vcx = pod<T>(cx), vcy = pod<T>(cy), vdx = pod<T>(dx), vdy = pod<T>(dy), vx = pod<T>(cenx), vy = pod<T>(ceny);
    cudaEventRecord(h->ev0, h->renderer.m_ComputeStream);
    const uint32_t rc = h->renderer.RenderPerturbBLAScaled<IterType, T>(a, &rt, &rf, vcx, vcy, vdx, vdy, vx, vy, (IterType)n, 1);
    cudaEventRecord(h->ev1, h->renderer.m_ComputeStream);
    h->renderer.SyncComputeStream();
    return rc;
}

template <class F> uint32_t by_type(int numeric, uint32_t iter_bytes, F &&f) {
    const bool u64 = iter_bytes == 8;
    switch (numeric) {
    case 0: return u64 ? f(uint64_t{}, float{}, float{}) : f(uint32_t{}, float{}, float{});
    case 1: return u64 ? f(uint64_t{}, double{}, double{}) : f(uint32_t{}, double{}, double{});
    case 2: return u64 ? f(uint64_t{}, CudaDblflt<MattDblflt>{}, CudaDblflt<MattDblflt>{}) : f(uint32_t{}, CudaDblflt<MattDblflt>{}, CudaDblflt<MattDblflt>{});
    case 5: return u64 ? f(uint64_t{}, HDRFloat<CudaDblflt<MattDblflt>>{}, CudaDblflt<MattDblflt>{}) : f(uint32_t{}, HDRFloat<CudaDblflt<MattDblflt>>{}, CudaDblflt<MattDblflt>{});
    case 3: return u64 ? f(uint64_t{}, HDRFloat<float>{}, float{}) : f(uint32_t{}, HDRFloat<float>{}, float{});
    case 4: return u64 ? f(uint64_t{}, HDRFloat<double>{}, double{}) : f(uint32_t{}, HDRFloat<double>{}, double{});
    default: return 10100;
    }
}

} // namespace

extern "C" {

void *refh_create() {
    Harness *h = new Harness();
    return h;
}
void refh_destroy(void *p) {
    Harness *h = (Harness *)p;
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
}
uint32_t refh_test_cuda() { return GPURenderer::TestCudaIsWorking(); }

uint32_t refh_init_memory(void *p, uint32_t iter_bytes, uint32_t w, uint32_t hgt, uint32_t aa, const void *pal,
                          uint32_t pal_iters, uint32_t aux, uint64_t gen, int reuse) {
    Harness *h = (Harness *)p;
    uint32_t rc = iter_bytes == 8
                      ? h->renderer.InitializeMemory<uint64_t>(w, hgt, aa, (const Color16 *)pal, pal_iters, aux, gen, reuse != 0)
                      : h->renderer.InitializeMemory<uint32_t>(w, hgt, aa, (const Color16 *)pal, pal_iters, aux, gen, reuse != 0);
    if (!h->ev0) { cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); }
    return rc;
}

uint32_t refh_init_perturb(void *p, uint32_t iter_bytes, int numeric, uint64_t gen, const void *orbit, uint64_t count,
                           uint64_t period, const void *xlow, const void *ylow, const void *las, uint64_t num_las,
                           const void *stages, uint64_t num_stages, const void *at, uint64_t stage_count, int use_at,
                           int is_valid) {
    Harness *h = (Harness *)p;
    return by_type(numeric, iter_bytes, [&](auto it, auto t, auto st) -> uint32_t {
        return init_perturb_t<decltype(it), decltype(t), decltype(st)>(h, gen, orbit, count, period, xlow, ylow, las,
                                                                        num_las, stages, num_stages, at, stage_count,
                                                                        use_at, is_valid);
    });
}

void refh_clear(void *p, uint32_t iter_bytes) {
    Harness *h = (Harness *)p;
    if (iter_bytes == 8) h->renderer.ClearMemory<uint64_t>(); else h->renderer.ClearMemory<uint32_t>();
}

uint32_t refh_init_perturb_rc(void *p, uint32_t iter_bytes, int numeric, uint64_t gen, const void *orbit, uint64_t count,
                              uint64_t full, uint64_t period, const void *xlow, const void *ylow, const void *las,
                              uint64_t num_las, const void *stages, uint64_t num_stages, const void *at,
                              uint64_t stage_count, int use_at, int is_valid) {
    Harness *h = (Harness *)p;
    return by_type(numeric, iter_bytes, [&](auto it, auto t, auto st) -> uint32_t {
        return init_perturb_rc_t<decltype(it), decltype(t), decltype(st)>(h, gen, orbit, count, full, period, xlow, ylow, las,
                                                                         num_las, stages, num_stages, at, stage_count,
                                                                         use_at, is_valid);
    });
}

uint32_t refh_render_lav2_rc(void *p, uint32_t iter_bytes, uint32_t alg, int numeric, int mode, const void *cx,
                             const void *cy, const void *dx, const void *dy, const void *cenx, const void *ceny,
                             uint64_t n) {
    Harness *h = (Harness *)p;
    cudaEventRecord(h->ev0, h->renderer.m_ComputeStream);
    const uint32_t rc = by_type(numeric, iter_bytes, [&](auto it, auto t, auto st) -> uint32_t {
        return render_lav2_rc_t<decltype(it), decltype(t), decltype(st)>(h, alg, mode, cx, cy, dx, dy, cenx, ceny, n);
    });
    cudaEventRecord(h->ev1, h->renderer.m_ComputeStream);
    return rc;
}

uint32_t refh_render_lav2(void *p, uint32_t iter_bytes, uint32_t alg, int numeric, int mode, const void *cx,
                          const void *cy, const void *dx, const void *dy, const void *cenx, const void *ceny,
                          uint64_t n) {
    Harness *h = (Harness *)p;
    cudaEventRecord(h->ev0, h->renderer.m_ComputeStream);
    const uint32_t rc = by_type(numeric, iter_bytes, [&](auto it, auto t, auto st) -> uint32_t {
        return render_lav2_t<decltype(it), decltype(t), decltype(st)>(h, alg, mode, cx, cy, dx, dy, cenx, ceny, n);
    });
    cudaEventRecord(h->ev1, h->renderer.m_ComputeStream);
    return rc;
}

// GPURenderer::RenderPerturbBLA<IterType,T>; ev0/ev1 bracket the per-call orbit + table upload and the kernel,
// exactly what the reference's own per-pixel timer sees (Fractal.cpp:2742)
uint32_t refh_render_bla(void *p, uint32_t iter_bytes, uint32_t alg, int numeric, const void *orbit, uint64_t count,
                         uint64_t period, const void *const *levels, const uint64_t *counts, uint32_t num_levels, int lm2,
                         const void *cx, const void *cy, const void *dx, const void *dy, const void *cenx,
                         const void *ceny, uint64_t n) {
    Harness *h = (Harness *)p;
    const bool u64 = iter_bytes == 8;
#define REFH_BLA(T)                                                                                                    \
    (u64 ? render_bla_t<uint64_t, T>(h, alg, orbit, count, period, levels, counts, num_levels, lm2, cx, cy, dx, dy, cenx, ceny, n) \
         : render_bla_t<uint32_t, T>(h, alg, orbit, count, period, levels, counts, num_levels, lm2, cx, cy, dx, dy, cenx, ceny, n))
    switch (numeric) {
    case 1: return REFH_BLA(double);
    case 3: return REFH_BLA(HDRFloat<float>);
    case 4: return REFH_BLA(HDRFloat<double>);
    default: return 10100;
    }
#undef REFH_BLA
}

uint32_t refh_render_scaled(void *p, uint32_t iter_bytes, uint32_t alg, int numeric, const void *orbit_t,
                            const void *orbit_f, uint64_t count, uint64_t period, const void *cx, const void *cy,
                            const void *dx, const void *dy, const void *cenx, const void *ceny, uint64_t n) {
    Harness *h = (Harness *)p;
    const bool u64 = iter_bytes == 8;
#define REFH_SC(T)                                                                                                     \
    (u64 ? render_scaled_t<uint64_t, T>(h, alg, orbit_t, orbit_f, count, period, cx, cy, dx, dy, cenx, ceny, n)        \
         : render_scaled_t<uint32_t, T>(h, alg, orbit_t, orbit_f, count, period, cx, cy, dx, dy, cenx, ceny, n))
    switch (numeric) {
    case 1: return REFH_SC(double);
    case 3: return REFH_SC(HDRFloat<float>);
    default: return 10100;
    }
#undef REFH_SC
}

uint32_t refh_render_direct(void *p, uint32_t iter_bytes, uint32_t alg, int numeric, const void *cx, const void *cy,
                            const void *dx, const void *dy, uint64_t n, int prec) {
    Harness *h = (Harness *)p;
    RenderAlgorithm a; *const_cast<RenderAlgorithmEnum *>(&a.Algorithm) = (RenderAlgorithmEnum)alg;
    cudaEventRecord(h->ev0, h->renderer.m_ComputeStream);
    uint32_t rc = 10100;
    if (numeric == 0) {
        rc = iter_bytes == 8 ? h->renderer.Render<uint64_t, float>(a, pod<float>(cx), pod<float>(cy), pod<float>(dx), pod<float>(dy), (uint64_t)n, prec)
                             : h->renderer.Render<uint32_t, float>(a, pod<float>(cx), pod<float>(cy), pod<float>(dx), pod<float>(dy), (uint32_t)n, prec);
    } else if (numeric == 1) {
        rc = iter_bytes == 8 ? h->renderer.Render<uint64_t, double>(a, pod<double>(cx), pod<double>(cy), pod<double>(dx), pod<double>(dy), (uint64_t)n, prec)
                             : h->renderer.Render<uint32_t, double>(a, pod<double>(cx), pod<double>(cy), pod<double>(dx), pod<double>(dy), (uint32_t)n, prec);
    }
#define REFH_DIRECT(TAG, T)                                                                                            \
    else if (numeric == TAG) {                                                                                         \
        rc = iter_bytes == 8 ? h->renderer.Render<uint64_t, T>(a, pod<T>(cx), pod<T>(cy), pod<T>(dx), pod<T>(dy), (uint64_t)n, prec) \
                             : h->renderer.Render<uint32_t, T>(a, pod<T>(cx), pod<T>(cy), pod<T>(dx), pod<T>(dy), (uint32_t)n, prec); \
    }
    REFH_DIRECT(2, MattDblflt)        // Gpu2x32
    REFH_DIRECT(6, MattDbldbl)        // Gpu2x64
    REFH_DIRECT(7, MattQFltflt)       // Gpu4x32
    REFH_DIRECT(8, MattQDbldbl)       // Gpu4x64
    REFH_DIRECT(4, HDRFloat<double>)  // GpuHDRx32 (converted to HDRFloat<CudaDblflt> inside Render)
#undef REFH_DIRECT
    cudaEventRecord(h->ev1, h->renderer.m_ComputeStream);
    return rc;
}

uint32_t refh_render_current(void *p, uint32_t iter_bytes, uint64_t n, void *iters, void *colors, void *red) {
    Harness *h = (Harness *)p;
    uint32_t rc = iter_bytes == 8
                      ? h->renderer.RenderCurrent<uint64_t>((uint64_t)n, (uint64_t *)iters, (Color16 *)colors, (ReductionResults *)red, false)
                      : h->renderer.RenderCurrent<uint32_t>((uint32_t)n, (uint32_t *)iters, (Color16 *)colors, (ReductionResults *)red, false);
    if (rc) return rc;
    return h->renderer.SyncComputeStream();
}

uint32_t refh_sync(void *p) { return ((Harness *)p)->renderer.SyncComputeStream(); }

// device time of the last render call (kernel only), ms
uint32_t refh_last_render_ms(void *p, float *ms) {
    Harness *h = (Harness *)p;
    cudaError_t e = cudaEventSynchronize(h->ev1);
    if (e != cudaSuccess) return e;
    return cudaEventElapsedTime(ms, h->ev0, h->ev1);
}

} // extern "C"
