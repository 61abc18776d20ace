// noize_stages.hpp — compiled-language host side over the C ABI (include/noize_b200.h).
//
// The reference's host side is C# (Unity); that toolchain is absent from the build image, so next to the
// C# shim (unity/Interop/NoizeB200.cs, not compilable here) this header mirrors the same stage API in C++:
// same class names, fields, defaults and error behaviour as the reference classes, one native call per stage.
//
//   PipelineStage / IStage            Pipeline/Stage/PipelineStage.cs:10-62, Pipeline/Interface.cs:7-45
//   StageIO, GeneratorData, MeshStageData   Pipeline/Stage/StageIO.cs:8-11, Pipeline/Stage/StageIOTypes/*.cs
//   PipelineWorkItem                  Pipeline/Stage/PipelineDefinition.cs:18-25
//   NoiseStage                        Noise/NoiseStage.cs:13-61
//   KernelFilterStage                 Filter/KernelFilterStage.cs:13-51
//   StageGaussianBlur/StageSmoothBlur Filter/Kernel/Blur/StageGaussianBlur.cs, StageSmoothBlur.cs
//   FlowMapStage                      Geologic/Stage/FlowMapStage.cs:16-221
//   MeshTileStage                     Mesh/Stage/MeshTileStage.cs:28-62
//   BasePipeline (scheduling chain)   Pipeline/Executable/Pipeline.cs:104-181
//   ErosionFilterStage                new: wraps ErosionKernelJob (Filter/Kernel/KernelJob.cs:317-350)
//
// A NativeSlice<float> is nz_slice_f32 {ptr, stride_bytes, length}.  Residency is owned by the stage objects: the first GPU
// stage of a work item opens a scope keyed by StageIO.uuid, the last one before a host consumer closes it (one D2H per
// dirty slice), exactly as the C# GpuStage does.  Header-only; link with -lnoize_b200.
#pragma once
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/noize_b200.h"

namespace noize {

struct NzError : std::runtime_error {
    int code;
    NzError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int32_t rc, const char* what) {
    if (rc < 0) throw NzError(rc, std::string(what) + ": " + nz_last_error());
}

// enums: same numeric order as the C# enums
enum class FractalNoise { Sin, Perlin, PeriodicPerlin, Simplex, RotatedSimplex, Cellular, DomainRotatedPerlin, DomainRotatedSimplex };
enum class KernelFilterType {
    Gauss9_S1, Gauss7_S1, Gauss5_S1, Gauss3_S1, Gauss9_S2, Gauss7_S2, Gauss5_S2, Gauss3_S2, Smooth3,
    Sobel3Horizontal, Sobel3Vertical, Sobel3_2D, Prewitt3Horizontal, Prewitt3Vertical
};
enum class MeshType { SquareGridHeightMap, OvershootSquareGridHeightMap };
enum class ConstantOperationType { MULTIPLY, BINARIZE };                          // Filter/ConstantStage.cs:15-18
enum class ReductionType { SUBTRACT, MULTIPLY, ROOTSUMSQUARES, MAX, MIN };        // Filter/Reduce/ReduceStage.cs:12-18

// ---- job handle + residency (same logic as unity/Interop/NoizeB200.cs: GpuResidency, CloseScopeJob, FlushScopeJob) ----
// The native calls of this mirror run inline where Unity would run an IJob on a worker thread, so a handle is complete
// when it is returned.
class JobHandle {
public:
    bool IsCompleted() const { return true; }
    void Complete() {}
};

// uuid -> residency scope (nz_scope_*) shared by the GPU stages that work on ONE work item (StageIO.uuid).
class GpuResidency {
    static std::map<std::string, int64_t>& scopes() {
        static std::map<std::string, int64_t> m;
        return m;
    }
    static std::mutex& mu() {
        static std::mutex m;
        return m;
    }

public:
    static int64_t Enter(const std::string& uuid) {
        std::lock_guard<std::mutex> lk(mu());
        auto it = scopes().find(uuid);
        if (it != scopes().end()) return it->second;
        const int64_t scope = nz_scope_create();
        if (scope < 0) check((int32_t)scope, "nz_scope_create");
        scopes()[uuid] = scope;
        return scope;
    }
    static void Close(const std::string& uuid) {              // CloseScopeJob
        int64_t scope = 0;
        {
            std::lock_guard<std::mutex> lk(mu());
            auto it = scopes().find(uuid);
            if (it == scopes().end()) return;
            scope = it->second;
            scopes().erase(it);
        }
        check(nz_scope_close(scope), "nz_scope_close");
    }
    static void Flush(const std::string& uuid, nz_slice_f32 data) {   // FlushScopeJob
        int64_t scope = 0;
        {
            std::lock_guard<std::mutex> lk(mu());
            auto it = scopes().find(uuid);
            if (it == scopes().end()) return;
            scope = it->second;
        }
        check(nz_scope_enter(scope), "nz_scope_enter");
        const int32_t rc = nz_flush_to_host(data.ptr);
        nz_scope_leave();
        check(rc, "nz_flush_to_host");
    }
    static bool IsOpen(const std::string& uuid) {
        std::lock_guard<std::mutex> lk(mu());
        return scopes().count(uuid) != 0;
    }
};

// ---- stage IO ---------------------------------------------------------------------------------------
struct StageIO {
    std::string uuid;
    nz_slice_f32 data{nullptr, 4, 0};
    virtual ~StageIO() = default;
};
struct GeneratorData : StageIO {
    int resolution = 512, xpos = 0, zpos = 0;
};
struct ReduceData : StageIO {                   // Pipeline/Stage/StageIOTypes/ReduceData.cs
    int resolution = 512, xpos = 0, zpos = 0;
    nz_slice_f32 rightData{nullptr, 4, 0};
};
struct DownsampleData : StageIO {               // Pipeline/Stage/StageIOTypes/DownsampleData.cs
    int resolution = 512, inputResolution = 512;
    nz_slice_f32 inputData{nullptr, 4, 0};
};
struct Mesh {                                   // Mesh.MeshData as PositionStream32.Setup declares it
    std::vector<nz_mesh_vertex> vertices;       // (R+1)^2 x 48 B
    std::vector<uint32_t> indices;              // 6 R^2
    float boundsCenter[3]{}, boundsSize[3]{};
};
struct MeshStageData : StageIO {
    int resolution = 512, inputResolution = 512, marginPix = 5;
    float tileSize = 512.f, tileHeight = 512.f;
    int xpos = 0, zpos = 0;
    Mesh* mesh = nullptr;
};
struct PipelineWorkItem {
    StageIO* data = nullptr;
    std::function<void(StageIO*)> completeAction;
    std::function<void(StageIO*, JobHandle)> scheduledAction;
    JobHandle dependency;
};

// ---- stage base -------------------------------------------------------------------------------------
class PipelineStage {
protected:
    JobHandle jobHandle;
    int dataLength = 0;

public:
    std::function<void(PipelineWorkItem&, JobHandle)> OnStageScheduledAction;
    PipelineStage* nextStage = nullptr;   // receiver of OnStageScheduledAction when it is a stage (C#: Delegate.Target)
    virtual bool IsGpuStage() const { return false; }
    virtual ~PipelineStage() = default;
    virtual void ResizeNativeContainers(int) {}
    virtual bool IsSchedulable(const PipelineWorkItem&) { return true; }
    template <class T>
    T* CheckRequirements(PipelineWorkItem& requirements) {
        T* d = dynamic_cast<T*>(requirements.data);
        if (!d) throw std::runtime_error("Unhandled stageio");
        if (d->data.length != dataLength) {
            dataLength = d->data.length;
            ResizeNativeContainers(dataLength);
        }
        return d;
    }
    virtual void Schedule(PipelineWorkItem& requirements, JobHandle dependency) = 0;
    void ReceiveHandledInput(PipelineWorkItem& requirements, JobHandle dependency) {
        Schedule(requirements, dependency);
        TransformData(requirements);
        OnStageScheduled(requirements, jobHandle);
    }
    virtual void TransformData(PipelineWorkItem&) {}
    virtual void OnStageScheduled(PipelineWorkItem& requirements, JobHandle) {
        if (OnStageScheduledAction) OnStageScheduledAction(requirements, jobHandle);
    }
    virtual void OnStageComplete() {}
    virtual void OnDestroy() {}
};

// Base of every stage that runs on the GPU: stage objects, not the pipeline, own residency (see NoizeB200.cs: GpuStage).
class GpuStage : public PipelineStage {
public:
    bool keepResident = false;   // last GPU stage of a pipeline: flush the result home, keep the tile in HBM
    bool IsGpuStage() const override { return true; }

protected:
    // NativeCallJob.Execute: enter the work item's scope, make the one blocking native call, leave
    template <class F>
    void Native(PipelineWorkItem& requirements, const char* what, F&& call) {
        const std::string& uuid = requirements.data->uuid;
        const int64_t scope = GpuResidency::Enter(uuid);
        check(nz_scope_enter(scope), "nz_scope_enter");
        const int32_t rc = call();
        nz_scope_leave();
        if (rc < 0) {
            std::string msg = std::string(what) + ": " + nz_last_error();
            try { GpuResidency::Close(uuid); } catch (...) {}
            throw NzError(rc, msg);
        }
        jobHandle = JobHandle();
    }

public:
    void OnStageScheduled(PipelineWorkItem& requirements, JobHandle dependency) override {
        // hand-over to something that reads HOST memory: the scope's close (or flush) goes in front of the handle
        if (!(nextStage && nextStage->IsGpuStage())) {
            if (keepResident) GpuResidency::Flush(requirements.data->uuid, requirements.data->data);
            else GpuResidency::Close(requirements.data->uuid);
        }
        PipelineStage::OnStageScheduled(requirements, dependency);
    }
};

class NoiseStage : public GpuStage {
public:
    FractalNoise noiseType = FractalNoise::Sin;
    float hurst = 0.f, startingAmplitude = 1.f, stepdown = 2.f, detuneRate = 0.f;
    int octaves = 1, noiseSize = 1000;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_fractal", [&] { return nz_fractal(d->data, d->resolution, (int)noiseType, hurst, startingAmplitude, stepdown, detuneRate, octaves,
                         d->xpos, d->zpos, noiseSize); });
    }
};

class KernelFilterStage : public GpuStage {
public:
    KernelFilterType filter = KernelFilterType::Gauss9_S1;
    int iterations = 1;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        // the reference chains `iterations` jobs; the GPU stage issues ONE fused call
        Native(requirements, "nz_kernel_filter", [&] { return nz_kernel_filter(d->data, nz_slice_f32{nullptr, 0, 0}, (int)filter, d->resolution, iterations); });
    }
};

class StageGaussianBlur : public GpuStage {
public:
    int iterations = 1, sigma = 0, width = 3;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_gauss_filter", [&] { return nz_gauss_filter(d->data, nz_slice_f32{nullptr, 0, 0}, width, sigma, d->resolution, iterations); });
    }
};

class StageSmoothBlur : public GpuStage {
public:
    int iterations = 1, width = 1;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_smooth_filter", [&] { return nz_smooth_filter(d->data, nz_slice_f32{nullptr, 0, 0}, width, d->resolution, iterations); });
    }
};

class ErosionFilterStage : public GpuStage {
public:
    int iterations = 5;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_min_erosion", [&] { return nz_min_erosion(d->data, d->resolution, iterations); });
    }
};

// ---- SURVEY section 8f rows ---------------------------------------------------------------------------------
class StageThermalErosion : public GpuStage {      // Filter/Kernel/Blur/StageThermalErosion.cs:13-29
public:
    int iterations = 1, talus = 45;
    float increment = 0.5f, meshHeightWidthRatio = 0.75f;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_thermal_erosion", [&] { return nz_thermal_erosion(d->data, (float)talus, increment, meshHeightWidthRatio, iterations, d->resolution); });
    }
};

class ErosionStageSubtractiveFlow : public GpuStage {   // Geologic/Stage/ErosionStageSubtractiveFlow.cs:17-247 (commented out upstream)
public:
    int flowIterations = 5;                              // serialised upstream, read by nothing (:19-20, :226-228)
    float normMin = -0.1f, normMax = 0.1f, erosiveFactor = 0.1f;
    int erosiveIterations = 5;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_subtractive_flow_erosion", [&] { return nz_subtractive_flow_erosion(d->data, d->resolution, erosiveIterations, erosiveFactor, normMin, normMax); });
    }
};

class ConstantStage : public GpuStage {            // Filter/ConstantStage.cs:13-60
public:
    ConstantOperationType operation = ConstantOperationType::MULTIPLY;
    float value = 0.5f;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_constant", [&] { return nz_constant(d->data, nz_slice_f32{nullptr, 0, 0}, (int)operation, value, d->resolution); });
    }
};

class ReduceStage : public GpuStage {              // Filter/Reduce/ReduceStage.cs:21-68
    GeneratorData transformed;

public:
    ReductionType operation = ReductionType::SUBTRACT;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        ReduceData* d = CheckRequirements<ReduceData>(requirements);
        Native(requirements, "nz_reduce", [&] { return nz_reduce(d->data, d->rightData, nz_slice_f32{nullptr, 0, 0}, (int)operation, d->resolution); });
    }
    void TransformData(PipelineWorkItem& inputData) override {   // downstream stages see a GeneratorData (:52-61)
        ReduceData* d = static_cast<ReduceData*>(inputData.data);
        transformed.uuid = d->uuid; transformed.data = d->data; transformed.resolution = d->resolution;
        transformed.xpos = d->xpos; transformed.zpos = d->zpos;
        inputData.data = &transformed;
    }
};

class CurveStage : public GpuStage {               // Filter/Curve/CurveStage.cs:13-73
    std::vector<float> curve;

public:
    std::function<float(float)> unityCurve = [](float t) { return t; };   // AnimationCurve.Evaluate
    int samples = 256;
    void ResizeNativeContainers(int) override {                           // ExtractCurve, :27-35
        curve.resize(samples);
        for (int i = 0; i < samples; i++) curve[i] = unityCurve((float)i / samples);
    }
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_curve", [&] { return nz_curve(d->data, nz_slice_f32{nullptr, 0, 0}, nz_slice_f32{curve.data(), 4, (int32_t)curve.size()}, d->resolution); });
    }
};

class CropStage : public GpuStage {                // Filter/Sample/CropStage.cs:13-19
public:
    bool center = false;   // false: the reference's behaviour (CropJob.Offset is never assigned: top-left corner)
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        DownsampleData* d = dynamic_cast<DownsampleData*>(requirements.data);
        if (!d) throw std::runtime_error("Unhandled stageio");
        Native(requirements, "nz_crop", [&] { return nz_crop(d->inputData, d->inputResolution, d->data, d->resolution, center ? (d->inputResolution - d->resolution) / 2 : 0); });
    }
};

class FlowMapStage : public GpuStage {
public:
    int iterations = 5;
    float normMin = -.1f, normMax = .1f;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        Native(requirements, "nz_flowmap", [&] { return nz_flowmap(d->data, d->resolution, iterations, normMin, normMax); });
    }
};

class MeshTileStage : public GpuStage {
public:
    MeshType meshType = MeshType::SquareGridHeightMap;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        MeshStageData* d = dynamic_cast<MeshStageData*>(requirements.data);
        if (!d || !d->mesh) throw std::runtime_error("Unhandled stageio");
        const int R = d->resolution;
        Mesh& m = *d->mesh;
        m.vertices.resize((size_t)(R + 1) * (R + 1));          // Mesh.AllocateWritableMeshData + PositionStream32.Setup
        m.indices.resize((size_t)6 * R * R);
        for (int k = 0; k < 3; k++) {
            m.boundsCenter[k] = 0.5f * (k == 1 ? d->tileHeight : d->tileSize);
            m.boundsSize[k] = k == 1 ? d->tileHeight : d->tileSize;
        }
        Native(requirements, "nz_heightmap_mesh", [&] { return nz_heightmap_mesh((int)meshType, m.vertices.data(), m.indices.data(), R, d->inputResolution, d->marginPix,
                                d->tileHeight, d->tileSize, d->data); });
    }
};

// ---- pipeline executor (scheduling chain of BasePipeline, Pipeline/Executable/Pipeline.cs:104-181) ------
class BasePipeline {
    std::vector<PipelineStage*> stages_;
    std::deque<PipelineWorkItem> queue_;
    bool running_ = false, queued_ = false;
    JobHandle handle_;
    PipelineWorkItem active_;

public:
    explicit BasePipeline(std::vector<PipelineStage*> stages) : stages_(std::move(stages)) {
        if (stages_.empty()) throw std::runtime_error("No stages in pipeline");
        for (size_t i = 0; i < stages_.size(); i++) {
            if (i + 1 < stages_.size()) {
                PipelineStage* next = stages_[i + 1];
                stages_[i]->nextStage = next;
                stages_[i]->OnStageScheduledAction = [next](PipelineWorkItem& w, JobHandle h) { next->ReceiveHandledInput(w, h); };
            } else {
                stages_[i]->OnStageScheduledAction = [this](PipelineWorkItem& w, JobHandle h) { OnPipelineFullyScheduled(w, h); };
            }
        }
    }
    void Enqueue(StageIO* input, std::function<void(StageIO*)> completeAction = nullptr) {
        PipelineWorkItem w;
        w.data = input;
        w.completeAction = std::move(completeAction);
        queue_.push_back(std::move(w));
    }
    void OnPipelineFullyScheduled(PipelineWorkItem& w, JobHandle h) {
        handle_ = h;
        queued_ = false;
        running_ = true;
        if (w.scheduledAction) w.scheduledAction(w.data, h);
    }
    void Update() {
        if (!running_ && !queued_ && !queue_.empty()) {
            active_ = std::move(queue_.front());
            queue_.pop_front();
            queued_ = true;
            stages_[0]->ReceiveHandledInput(active_, active_.dependency);
        }
    }
    void LateUpdate() {
        if (running_) {
            handle_.Complete();
            running_ = false;
            for (auto* s : stages_) s->OnStageComplete();
            if (active_.completeAction) active_.completeAction(active_.data);
        }
    }
    void Run(StageIO* input, std::function<void(StageIO*)> completeAction = nullptr) {
        Enqueue(input, std::move(completeAction));
        Update();
        LateUpdate();
    }
};

}  // namespace noize
