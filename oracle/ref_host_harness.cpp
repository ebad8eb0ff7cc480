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

#include <atomic>
#include <cstring>
#include <deque>
#include <memory>
#include <thread>
#include <vector>

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

// The reference's CPU renderer for HDRx32 + LAv2, Cpu32PerturbedBLAV2HDR: the per-pixel loop of
// Fractal::CalcCpuPerturbationFractalLAV2<IterType, float, Disable> (Fractal.cpp:2485-2691) with its row claiming
// (:2523-2543) and hardware_concurrency() threads (:2684-2690).  Fractal.cpp itself cannot be compiled here (OpenGL
// headers), so the loop is restated below, statement for statement, on the reference's OWN types and compiled methods
// (HDRFloat / HDRFloatComplex operators, ATInfo::PerformAT, LAReference::getLA / isLAStageInvalid, LAstep::Evaluate /
// getZ, PerturbationResults::GetComplex): every arithmetic operation executed is the reference's code.  Used as the
// timed CPU baseline only -- its iteration counts differ from the GPU algorithms' by design (bailout 256, the opposite
// sense of isLAStageInvalid: LAReference.cpp:1076-1081 vs GPU_LAReference.h:241-255).
// row_step / col_step > 1 render a regular sub-grid (a bounded sample of the frame); returns the sum of the iteration
// counts of the pixels rendered.  out (IterType[h][w]) may be NULL.
uint64_t refhost_cpu_lav2(void *handle, int iter_bytes, int w, int h, const void *dx_p, const void *dy_p, const void *cx_p,
                          const void *cy_p, uint64_t n_iterations, void *out, int row_step, int col_step, int n_threads) {
    auto run = [&](auto it) -> uint64_t {
        using IterType = decltype(it);
        using SubType = float;
        using T = HDRFloat<SubType>;
        using TComplex = HDRFloatComplex<SubType>;
        constexpr PerturbExtras PExtras = PerturbExtras::Disable;
        auto *hd = static_cast<LaHolder<IterType> *>(handle);
        auto *results = hd->results.get();
        auto &LaReference = *hd->la;
        T dx, dy, centerX, centerY;
        std::memcpy(&dx, dx_p, sizeof(T)); std::memcpy(&dy, dy_p, sizeof(T));
        std::memcpy(&centerX, cx_p, sizeof(T)); std::memcpy(&centerY, cy_p, sizeof(T));
        HdrReduce(dx); HdrReduce(dy); HdrReduce(centerX); HdrReduce(centerY);
        const IterType NumIterations = (IterType)n_iterations;
        const size_t num_threads = n_threads > 0 ? (size_t)n_threads : std::thread::hardware_concurrency();
        std::deque<std::atomic_uint64_t> atomics;
        atomics.resize(h);
        std::atomic<uint64_t> total{0};
        auto one_thread = [&]() {
            auto compressionHelper{std::make_unique<RuntimeDecompressor<IterType, T, PExtras>>(*results)};
            uint64_t local = 0;
            for (size_t y = 0; y < (size_t)h; y += row_step) {
                if (atomics[y] != 0) continue;
                uint64_t expected = 0;
                if (atomics[y].compare_exchange_strong(expected, 1llu) == false) continue;
                for (size_t x = 0; x < (size_t)w; x += col_step) {
                    IterType BLA2SkippedIterations = 0;
                    TComplex DeltaSub0;
                    TComplex DeltaSubN;
                    T deltaReal = dx * (SubType)x;
                    HdrReduce(deltaReal);
                    deltaReal -= centerX;
                    T deltaImaginary = -dy * (SubType)y;
                    HdrReduce(deltaImaginary);
                    deltaImaginary -= centerY;
                    HdrReduce(deltaReal);
                    HdrReduce(deltaImaginary);
                    DeltaSub0 = {deltaReal, deltaImaginary};
                    DeltaSubN = {0, 0};
                    if (LaReference.IsValid() && LaReference.UseAT() && LaReference.GetAT().isValid(DeltaSub0)) {
                        ATResult<IterType, T, SubType> res;
                        LaReference.GetAT().PerformAT(NumIterations, DeltaSub0, res);
                        BLA2SkippedIterations = res.bla_iterations;
                        DeltaSubN = res.dz;
                    }
                    IterType iterations = 0;
                    IterType RefIteration = 0;
                    IterType MaxRefIteration = (IterType)results->GetCountOrbitEntries() - 1;
                    iterations = BLA2SkippedIterations;
                    TComplex complex0{deltaReal, deltaImaginary};
                    if (iterations != 0 && RefIteration < MaxRefIteration) {
                        complex0 = results->template GetComplex<SubType>(*compressionHelper, RefIteration) + DeltaSubN;
                    } else if (iterations != 0 && results->GetPeriodMaybeZero() != 0) {
                        RefIteration = RefIteration % results->GetPeriodMaybeZero();
                        complex0 = results->template GetComplex<SubType>(*compressionHelper, RefIteration) + DeltaSubN;
                    }
                    auto CurrentLAStage = LaReference.IsValid() ? LaReference.GetLAStageCount() : 0;
                    while (CurrentLAStage > 0) {
                        CurrentLAStage--;
                        auto LAIndex = LaReference.getLAIndex(CurrentLAStage);
                        if (LaReference.isLAStageInvalid(LAIndex, DeltaSub0)) continue;
                        auto MacroItCount = LaReference.getMacroItCount(CurrentLAStage);
                        auto j = RefIteration;
                        while (iterations < NumIterations) {
                            auto las = LaReference.getLA(LAIndex, DeltaSubN, (IterType)j, (IterType)iterations, NumIterations);
                            if (las.unusable) {
                                RefIteration = las.nextStageLAindex;
                                break;
                            }
                            iterations += las.step;
                            DeltaSubN = las.Evaluate(DeltaSub0);
                            complex0 = las.getZ(DeltaSubN);
                            j++;
                            auto lhs = complex0.chebychevNorm();
                            HdrReduce(lhs);
                            auto rhs = DeltaSubN.chebychevNorm();
                            HdrReduce(rhs);
                            if (HdrCompareToBothPositiveReducedLT(lhs, rhs) || j >= MacroItCount) {
                                DeltaSubN = complex0;
                                j = 0;
                            }
                        }
                        if (iterations >= NumIterations) break;
                    }
                    T normSquared{};
                    if (iterations < NumIterations) normSquared = complex0.norm_squared();
                    for (; iterations < NumIterations; iterations++) {
                        auto curIter = results->template GetComplex<SubType>(*compressionHelper, RefIteration);
                        curIter = curIter * T(2);
                        curIter = curIter + DeltaSubN;
                        DeltaSubN = DeltaSubN * curIter;
                        DeltaSubN = DeltaSubN + DeltaSub0;
                        HdrReduce(DeltaSubN);
                        RefIteration++;
                        complex0 = results->template GetComplex<SubType>(*compressionHelper, RefIteration) + DeltaSubN;
                        HdrReduce(complex0);
                        normSquared = complex0.norm_squared();
                        HdrReduce(normSquared);
                        auto DeltaNormSquared = DeltaSubN.norm_squared();
                        HdrReduce(DeltaNormSquared);
                        if (HdrCompareToBothPositiveReducedGT(normSquared, T(256))) break;
                        if (HdrCompareToBothPositiveReducedLT(normSquared, DeltaNormSquared) || (RefIteration >= MaxRefIteration)) {
                            DeltaSubN = complex0;
                            RefIteration = 0;
                        }
                    }
                    if (out) static_cast<IterType *>(out)[y * (size_t)w + x] = static_cast<IterType>(iterations);
                    local += (uint64_t)iterations;
                }
            }
            total += local;
        };
        std::vector<std::unique_ptr<std::thread>> threads;
        threads.reserve(num_threads);
        for (size_t t = 0; t < num_threads; t++) threads.push_back(std::make_unique<std::thread>(one_thread));
        for (auto &t : threads) t->join();
        return total.load();
    };
    return iter_bytes == 8 ? run(uint64_t{}) : run(uint32_t{});
}

void refhost_free(void *handle, int iter_bytes) {
    if (iter_bytes == 8) delete static_cast<LaHolder<uint64_t> *>(handle);
    else delete static_cast<LaHolder<uint32_t> *>(handle);
}

} // extern "C"
