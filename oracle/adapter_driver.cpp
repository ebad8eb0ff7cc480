// adapter_driver.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref/adapter_driver, built by oracle/Makefile where the reference
// tree is mounted; run by tests/test_adapter.py on the GPU box).
//
// A C++ caller of include/fs_gpu_adapter.hpp written the way Fractal.cpp calls `GPURenderer`
// (CalcGpuPerturbationFractalLAv2 Fractal.cpp:2760-2850, CalcGpuPerturbationFractalBLA :2693-2758, CalcGpuFractal
// :1894-1915, result pull :1516-1532): the reference's OWN host objects -- PerturbationResults filled with an orbit,
// LAReference::GenerateApproximationData, BLAS::Init, GPUPerturbResults -- are handed to the adapter class, which forwards
// them to libfsgpu.so.  The iteration buffer goes to a file the test compares with the reference-kernel fixtures.
//
// usage: adapter_driver <case.bin> <out.bin>      case.bin (little endian):
//   u32 kind (1 = HDRx32 LAv2 Full, 2 = HDRx32 BLA, 3 = direct f64), u32 w, u32 h, u64 n_iterations, u64 orbit_count,
//   u64 period, then for kinds 1/2: HDRFloat<float> dx, dy, centerX, centerY, radius (8 bytes each) and orbit_count x 16
//   bytes of GPUReferenceIter<HDRFloat<float>>; for kind 3: double cx, cy, dx, dy.
#include "stdafx.h"

#include "BLA.h"
#include "BLAS.h"
#include "LAstep.h"
#include "GPU_Types.h"
#include "LAParameters.h"
#include "LAReference.h"
#include "PerturbationResults.h"
#include "RefOrbitCalc.h"
#include "RenderAlgorithm.h"

#define FS_GPU_ADAPTER_CLASS GPURenderer
#include "fs_gpu_adapter.hpp"

#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

namespace {

template <class V> bool read_pod(FILE *f, V &v) { return std::fread(&v, sizeof(V), 1, f) == 1; }

int fail(const char *what, uint32_t rc) {
    std::fprintf(stderr, "adapter_driver: %s failed: %u (%s)\n", what, rc, GPURenderer::ConvertErrorToString(rc));
    return 2;
}

// Never called: instantiates every member template of the adapter with the argument types the reference instantiates
// (GPU_Render.cu:227-230, 409-429, 511-537, 583-594, 849-991, 1204-1300, 1380-1436, 1610-1692, 1807-1818), so that a
// signature drifting from GPU_Render.h is a compile error here.
template <typename IterType> void instantiate_everything(GPURenderer &r) {
    const RenderAlgorithm alg = GetRenderAlgorithmTupleEntry(RenderAlgorithmEnum::AUTO);
    const IterType n = 1;
    using HF = HDRFloat<float>;
    using HD = HDRFloat<double>;
    using C2 = CudaDblflt<MattDblflt>;
    using H2 = HDRFloat<CudaDblflt<MattDblflt>>;
    r.template InitializeMemory<IterType>(16, 8, 1, nullptr, 0, 0, 0, false);
    r.template ClearMemory<IterType>();
    r.template Render<IterType, float>(alg, 0.f, 0.f, 0.f, 0.f, n, 1);
    r.template Render<IterType, double>(alg, 0., 0., 0., 0., n, 1);
    r.template Render<IterType, MattDblflt>(alg, MattDblflt{}, MattDblflt{}, MattDblflt{}, MattDblflt{}, n, 1);
    r.template Render<IterType, MattDbldbl>(alg, MattDbldbl{}, MattDbldbl{}, MattDbldbl{}, MattDbldbl{}, n, 1);
    r.template Render<IterType, MattQFltflt>(alg, MattQFltflt{}, MattQFltflt{}, MattQFltflt{}, MattQFltflt{}, n, 1);
    r.template Render<IterType, MattQDbldbl>(alg, MattQDbldbl{}, MattQDbldbl{}, MattQDbldbl{}, MattQDbldbl{}, n, 1);
    r.template Render<IterType, HD>(alg, HD{}, HD{}, HD{}, HD{}, n, 1);
    r.template RenderPerturbBLAScaled<IterType, double>(alg, nullptr, nullptr, 0., 0., 0., 0., 0., 0., n, 1);
    r.template RenderPerturbBLAScaled<IterType, HF>(alg, nullptr, nullptr, HF{}, HF{}, HF{}, HF{}, HF{}, HF{}, n, 1);
    r.template RenderPerturbBLA<IterType, double>(alg, nullptr, nullptr, 0., 0., 0., 0., 0., 0., n, 1);
    r.template RenderPerturbBLA<IterType, HF>(alg, nullptr, nullptr, HF{}, HF{}, HF{}, HF{}, HF{}, HF{}, n, 1);
    r.template RenderPerturbBLA<IterType, HD>(alg, nullptr, nullptr, HD{}, HD{}, HD{}, HD{}, HD{}, HD{}, n, 1);
#define FS_LAV2(T, Sub, Px)                                                                                             \
    r.template InitializePerturb<IterType, T, Sub, Px, T>(0, nullptr, 0, nullptr, nullptr);                             \
    r.template RenderPerturbLAv2<IterType, T, Sub, LAv2Mode::Full, Px>(alg, T{}, T{}, T{}, T{}, T{}, T{}, n);           \
    r.template RenderPerturbLAv2<IterType, T, Sub, LAv2Mode::PO, Px>(alg, T{}, T{}, T{}, T{}, T{}, T{}, n);             \
    r.template RenderPerturbLAv2<IterType, T, Sub, LAv2Mode::LAO, Px>(alg, T{}, T{}, T{}, T{}, T{}, T{}, n);
    FS_LAV2(float, float, PerturbExtras::Disable)
    FS_LAV2(double, double, PerturbExtras::Disable)
    FS_LAV2(HF, float, PerturbExtras::Disable)
    FS_LAV2(HD, double, PerturbExtras::Disable)
    FS_LAV2(float, float, PerturbExtras::SimpleCompression)
    FS_LAV2(double, double, PerturbExtras::SimpleCompression)
    FS_LAV2(HF, float, PerturbExtras::SimpleCompression)
    FS_LAV2(HD, double, PerturbExtras::SimpleCompression)
#undef FS_LAV2
    // the 2x32 types have no host LAReference instantiation of their own (LAReference.cpp:1032-1048): the reference
    // converts a double-based table element-wise (RefOrbitCalc.cpp:2490-2516); only the render entry is typed by them
    r.template RenderPerturbLAv2<IterType, C2, C2, LAv2Mode::Full, PerturbExtras::Disable>(alg, C2{}, C2{}, C2{}, C2{}, C2{}, C2{}, n);
    r.template RenderPerturbLAv2<IterType, H2, C2, LAv2Mode::Full, PerturbExtras::Disable>(alg, H2{}, H2{}, H2{}, H2{}, H2{}, H2{}, n);
    IterType *iters = nullptr;
    r.template RenderCurrent<IterType>(n, iters, nullptr, nullptr, true);
    r.SyncComputeStream(); r.SyncDisplayStream(); r.QueryComputeStream(); r.EnqueueComputeDoneCallback();
    r.ResetComputeDoneFlag(); r.IsComputeDone(); r.SetComputeDoneNotification(nullptr, nullptr);
    (void)r.GetWidth(); (void)r.GetHeight();
}

} // namespace

int main(int argc, char **argv) {
    if (argc == 2 && std::strcmp(argv[1], "--never") == 0) { // keeps the instantiations alive without running them
        GPURenderer r;
        instantiate_everything<uint32_t>(r);
        instantiate_everything<uint64_t>(r);
    }
    if (argc != 3) {
        std::fprintf(stderr, "usage: adapter_driver <case.bin> <out.bin>\n");
        return 1;
    }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return fail("open case", 1);
    uint32_t kind = 0, w = 0, h = 0;
    uint64_t n_iter = 0, count = 0, period = 0;
    if (!read_pod(f, kind) || !read_pod(f, w) || !read_pod(f, h) || !read_pod(f, n_iter) || !read_pod(f, count) || !read_pod(f, period))
        return fail("read header", 1);
    using IterType = uint32_t;
    using T = HDRFloat<float>;
    if (!GPURenderer::TestCudaIsWorking()) return fail("TestCudaIsWorking", 0);
    GPURenderer renderer;
    std::vector<Color16> palette(256);
    for (size_t i = 0; i < palette.size(); i++) palette[i] = Color16{(uint16_t)(i * 257), (uint16_t)(65535 - i * 257), (uint16_t)(i * 131), 65535};
    uint32_t rc = renderer.InitializeMemory<IterType>(w, h, 1, palette.data(), (uint32_t)palette.size(), 0, 1, false);
    if (rc) return fail("InitializeMemory", rc);
    const size_t wp = (w + 15) / 16 * 16, hp = (h + 7) / 8 * 8;
    std::vector<IterType> iters(wp * hp, 0xDEADBEEFu);
    ReductionResults red{};
    if (kind == 3) {
        double c[4];
        if (std::fread(c, sizeof(double), 4, f) != 4) return fail("read coords", 1);
        renderer.ClearMemory<IterType>();
        rc = renderer.Render<IterType, double>(GetRenderAlgorithmTupleEntry(RenderAlgorithmEnum::Gpu1x64), c[0], c[1], c[2], c[3],
                                               (IterType)n_iter, 1);
        if (rc) return fail("Render", rc);
    } else {
        T dx, dy, centerX, centerY, radius;
        if (!read_pod(f, dx) || !read_pod(f, dy) || !read_pod(f, centerX) || !read_pod(f, centerY) || !read_pod(f, radius))
            return fail("read coords", 1);
        std::vector<GPUReferenceIter<T, PerturbExtras::Disable>> orbit(count);
        if (std::fread(orbit.data(), sizeof(orbit[0]), count, f) != count) return fail("read orbit", 1);
        // the reference's own host objects
        auto results = std::make_unique<PerturbationResults<IterType, T, PerturbExtras::Disable>>(AddPointOptions::DontSave, 1);
        const HighPrecision zero{0};
        results->InitResults(RefOrbitCalc::ReuseMode::DontSaveForReuse, zero, zero, radius, (IterType)n_iter, (size_t)count + 16);
        for (uint64_t i = 1; i < count; i++) results->AddUncompressedIteration(orbit[i]);
        results->SetPeriodMaybeZero((IterType)period);
        GPUPerturbResults<IterType, T, PerturbExtras::Disable> gpu_results{
            (IterType)results->GetCountOrbitEntries(), (IterType)results->GetCountOrbitEntries(), results->GetOrbitXLow(),
            results->GetOrbitYLow(), results->GetOrbitData(), results->GetPeriodMaybeZero()};
        if (kind == 1) {
            LAParameters params;
            auto la = std::make_unique<LAReference<IterType, T, float, PerturbExtras::Disable>>(params, AddPointOptions::DontSave, L"", L"");
            la->GenerateApproximationData(*results, results->GetMaxRadius(), false);
            rc = renderer.InitializePerturb<IterType, T, float, PerturbExtras::Disable, T>(results->GetGenerationNumber(), &gpu_results, 0,
                                                                                           nullptr, la.get());
            if (rc) return fail("InitializePerturb", rc);
            renderer.ClearMemory<IterType>();
            rc = renderer.RenderPerturbLAv2<IterType, T, float, LAv2Mode::Full, PerturbExtras::Disable>(
                GetRenderAlgorithmTupleEntry(RenderAlgorithmEnum::GpuHDRx32PerturbedLAv2), T{}, T{}, dx, dy, centerX, centerY, (IterType)n_iter);
            if (rc) return fail("RenderPerturbLAv2", rc);
        } else {
            BLAS<IterType, T> blas(*results);
            blas.Init(results->GetCountOrbitEntries(), results->GetMaxRadius());
            renderer.ClearMemory<IterType>();
            rc = renderer.RenderPerturbBLA<IterType, T>(GetRenderAlgorithmTupleEntry(RenderAlgorithmEnum::GpuHDRx32PerturbedBLA), &gpu_results,
                                                        &blas, T{}, T{}, dx, dy, centerX, centerY, (IterType)n_iter, 1);
            if (rc) return fail("RenderPerturbBLA", rc);
        }
    }
    std::fclose(f);
    // Fractal.cpp:1516-1532: sync, pull, sync
    if ((rc = renderer.SyncComputeStream())) return fail("SyncComputeStream", rc);
    if ((rc = renderer.RenderCurrent<IterType>((IterType)n_iter, iters.data(), nullptr, &red, false))) return fail("RenderCurrent", rc);
    if ((rc = renderer.SyncComputeStream())) return fail("SyncComputeStream", rc);
    FILE *o = std::fopen(argv[2], "wb");
    if (!o) return fail("open output", 1);
    const uint64_t hdr[5] = {wp, hp, red.Min, red.Max, red.Sum};
    std::fwrite(hdr, sizeof(hdr), 1, o);
    std::fwrite(iters.data(), sizeof(IterType), iters.size(), o);
    std::fclose(o);
    std::printf("adapter_driver: kind %u %ux%u n=%llu sum=%llu min=%llu max=%llu\n", kind, w, h, (unsigned long long)n_iter,
                (unsigned long long)red.Sum, (unsigned long long)red.Min, (unsigned long long)red.Max);
    return 0;
}
