// ref_host_harness.cpp -- TEST INFRASTRUCTURE ONLY (built into oracle/_ref/libref_host.so by oracle/Makefile, loaded by tests/).
//
// Flat C entry points around the REFERENCE's own host-side table builders, compiled from the sources where they lie
// under /root/reference (nothing is copied):
//   LAReference<IterType, HDRFloat<float>, float, Disable>::GenerateApproximationData   FractalSharkLib/LAReference.cpp:971-1017
//   (CreateLAFromOrbit :28-207, CreateLAFromOrbitMT :215-771, CreateNewLAStage :774-968, CreateATFromLA :1050-1074,
//    LAInfoDeep Step/Composite/CreateAT  HpSharkFloatLib/LAInfoDeep.h:109-502)
//   BLAS<IterType, HDRFloat<float>>::Init                                                 FractalSharkLib/BLAS.cpp:212-254
// The orbit is handed in by the caller (the in-tree generator's), so what is compared is table construction alone.
#include "stdafx.h"

#include "BLAS.h"
#include "LAParameters.h"
#include "LAReference.h"
#include "PerturbationResults.h"
#include "RefOrbitCalc.h"

#include <cstring>
#include <memory>

namespace {

template <class IterT> struct LaHolder {
    using T = HDRFloat<float>;
    std::unique_ptr<PerturbationResults<IterT, T, PerturbExtras::Disable>> results;
    std::unique_ptr<LAReference<IterT, T, float, PerturbExtras::Disable>> la;
    std::unique_ptr<BLAS<IterT, T, PerturbExtras::Disable>> blas;
};

template <class IterT>
LaHolder<IterT> *build_results(const void *orbit_elems, uint64_t count, const void *radius_hdr, uint64_t n_iterations) {
    using T = HDRFloat<float>;
    auto *h = new LaHolder<IterT>();
    h->results = std::make_unique<PerturbationResults<IterT, T, PerturbExtras::Disable>>(AddPointOptions::DontSave, 1);
    T radius;
    std::memcpy(&radius, radius_hdr, sizeof(T));
    const HighPrecision zero{0};
    // InitResults pushes the all-zero element 0 itself (PerturbationResults.cpp:831-868)
    h->results->InitResults(RefOrbitCalc::ReuseMode::DontSaveForReuse, zero, zero, radius, (IterT)n_iterations, (size_t)count + 16);
    const auto *e = static_cast<const GPUReferenceIter<T, PerturbExtras::Disable> *>(orbit_elems);
    for (uint64_t i = 1; i < count; i++) h->results->AddUncompressedIteration(e[i]);
    return h;
}

} // namespace

extern "C" {

// iter_bytes 4|8.  threading: 0 = the reference's default (LAParameters.h:66-75), 1 = single-threaded, 2 = multi-threaded.
void *refhost_build_la(int iter_bytes, const void *orbit_elems, uint64_t count, const void *radius_hdr, uint64_t n_iterations,
                       int threading) {
    auto run = [&](auto it) -> void * {
        using IterT = decltype(it);
        using T = HDRFloat<float>;
        auto *h = build_results<IterT>(orbit_elems, count, radius_hdr, n_iterations);
        LAParameters params;
        if (threading == 1) params.SetThreading(LAParameters::LAThreadingAlgorithm::SingleThreaded);
        if (threading == 2) params.SetThreading(LAParameters::LAThreadingAlgorithm::MultiThreaded);
        h->la = std::make_unique<LAReference<IterT, T, float, PerturbExtras::Disable>>(params, AddPointOptions::DontSave, L"", L"");
        T radius;
        std::memcpy(&radius, radius_hdr, sizeof(T));
        h->la->GenerateApproximationData(*h->results, radius, false);
        return h;
    };
    return iter_bytes == 8 ? run(uint64_t{}) : run(uint32_t{});
}

// out[0] = NumLAs, [1] = NumStages (trimmed), [2] = LAStageCount, [3] = UseAT, [4] = IsValid, [5] = sizeof(LAInfoDeep),
// [6] = sizeof(ATInfo), [7] = sizeof(LAStageInfo)
void refhost_la_info(void *handle, int iter_bytes, uint64_t *out) {
    auto run = [&](auto it) {
        using IterT = decltype(it);
        auto *h = static_cast<LaHolder<IterT> *>(handle);
        out[0] = h->la->GetLAs().GetSize();
        out[1] = h->la->GetLAStages().GetSize();
        out[2] = h->la->GetLAStageCount();
        out[3] = h->la->UseAT();
        out[4] = h->la->IsValid();
        out[5] = sizeof(LAInfoDeep<IterT, HDRFloat<float>, float, PerturbExtras::Disable>);
        out[6] = sizeof(ATInfo<IterT, HDRFloat<float>, float>);
        out[7] = sizeof(LAStageInfo<IterT>);
    };
    if (iter_bytes == 8) run(uint64_t{}); else run(uint32_t{});
}

void refhost_la_copy(void *handle, int iter_bytes, void *las, void *stages, void *at) {
    auto run = [&](auto it) {
        using IterT = decltype(it);
        auto *h = static_cast<LaHolder<IterT> *>(handle);
        const auto &L = h->la->GetLAs();
        const auto &S = h->la->GetLAStages();
        if (las && L.GetSize()) std::memcpy(las, L.GetData(), L.GetSize() * sizeof(LAInfoDeep<IterT, HDRFloat<float>, float, PerturbExtras::Disable>));
        if (stages && S.GetSize()) std::memcpy(stages, S.GetData(), S.GetSize() * sizeof(LAStageInfo<IterT>));
        if (at) std::memcpy(at, &h->la->GetAT(), sizeof(ATInfo<IterT, HDRFloat<float>, float>));
    };
    if (iter_bytes == 8) run(uint64_t{}); else run(uint32_t{});
}

// BLAS<IterType, HDRFloat<float>>::Init(GetCountOrbitEntries(), GetMaxRadius())  (Fractal.cpp:2739-2740, BLAS.cpp:212-254)
// on the orbit of a handle made by refhost_build_la.  Returns the number of levels (m_B.size()); out_lm2 = m_LM2.
uint64_t refhost_build_blas(void *handle, int iter_bytes, int32_t *out_lm2) {
    auto run = [&](auto it) -> uint64_t {
        using IterT = decltype(it);
        auto *h = static_cast<LaHolder<IterT> *>(handle);
        h->blas = std::make_unique<BLAS<IterT, HDRFloat<float>, PerturbExtras::Disable>>(*h->results);
        h->blas->Init(h->results->GetCountOrbitEntries(), h->results->GetMaxRadius());
        *out_lm2 = h->blas->m_LM2;
        return h->blas->m_B.size();
    };
    return iter_bytes == 8 ? run(uint64_t{}) : run(uint32_t{});
}
// element count of level `level`; when `out` is given the BLA<HDRFloat<float>> records (44 bytes each) are copied there
uint64_t refhost_blas_level(void *handle, int iter_bytes, uint64_t level, void *out) {
    auto run = [&](auto it) -> uint64_t {
        using IterT = decltype(it);
        auto *h = static_cast<LaHolder<IterT> *>(handle);
        const auto &v = h->blas->m_B[level];
        static_assert(sizeof(BLA<HDRFloat<float>>) == 44, "BLA<HDRFloat<float>> layout (BLA.h:7-14)");
        if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(BLA<HDRFloat<float>>));
        return v.size();
    };
    return iter_bytes == 8 ? run(uint64_t{}) : run(uint32_t{});
}

void refhost_free(void *handle, int iter_bytes) {
    if (iter_bytes == 8) delete static_cast<LaHolder<uint64_t> *>(handle);
    else delete static_cast<LaHolder<uint32_t> *>(handle);
}

} // extern "C"
