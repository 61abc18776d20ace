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
// A NativeSlice<float> is nz_slice_f32 {ptr, stride_bytes, length}.  JobHandle::Complete() returns when the
// host buffers hold the results (the reference handle's contract); until then the device copies stay
// resident (nz_pipeline_begin / nz_pipeline_end).  Header-only; link with -lnoize_b200.
#pragma once
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
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

// ---- job handle: one residency scope shared by the stages of a scheduled chain ----------------------
class JobHandle {
    struct Scope {
        bool open = true;
        Scope() { check(nz_pipeline_begin(), "nz_pipeline_begin"); }
        void close() {
            if (open) {
                open = false;
                check(nz_pipeline_end(), "nz_pipeline_end");
            }
        }
        ~Scope() {
            if (open) nz_pipeline_end();
        }
    };
    std::shared_ptr<Scope> scope_;

public:
    JobHandle() = default;
    static JobHandle Chain(const JobHandle& dependency) {
        if (dependency.scope_ && dependency.scope_->open) return dependency;
        JobHandle h;
        h.scope_ = std::make_shared<Scope>();
        return h;
    }
    bool IsCompleted() const { return !scope_ || !scope_->open; }
    void Complete() {
        if (scope_) scope_->close();
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

class NoiseStage : public PipelineStage {
public:
    FractalNoise noiseType = FractalNoise::Sin;
    float hurst = 0.f, startingAmplitude = 1.f, stepdown = 2.f, detuneRate = 0.f;
    int octaves = 1, noiseSize = 1000;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_fractal(d->data, d->resolution, (int)noiseType, hurst, startingAmplitude, stepdown, detuneRate, octaves,
                         d->xpos, d->zpos, noiseSize), "nz_fractal");
    }
};

class KernelFilterStage : public PipelineStage {
public:
    KernelFilterType filter = KernelFilterType::Gauss9_S1;
    int iterations = 1;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        // the reference chains `iterations` jobs; the GPU stage issues ONE fused call
        check(nz_kernel_filter(d->data, nz_slice_f32{nullptr, 0, 0}, (int)filter, d->resolution, iterations), "nz_kernel_filter");
    }
};

class StageGaussianBlur : public PipelineStage {
public:
    int iterations = 1, sigma = 0, width = 3;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_gauss_filter(d->data, nz_slice_f32{nullptr, 0, 0}, width, sigma, d->resolution, iterations), "nz_gauss_filter");
    }
};

class StageSmoothBlur : public PipelineStage {
public:
    int iterations = 1, width = 1;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_smooth_filter(d->data, nz_slice_f32{nullptr, 0, 0}, width, d->resolution, iterations), "nz_smooth_filter");
    }
};

class ErosionFilterStage : public PipelineStage {
public:
    int iterations = 5;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_min_erosion(d->data, d->resolution, iterations), "nz_min_erosion");
    }
};

// ---- SURVEY section 8f rows ---------------------------------------------------------------------------------
class StageThermalErosion : public PipelineStage {      // Filter/Kernel/Blur/StageThermalErosion.cs:13-29
public:
    int iterations = 1, talus = 45;
    float increment = 0.5f, meshHeightWidthRatio = 0.75f;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_thermal_erosion(d->data, (float)talus, increment, meshHeightWidthRatio, iterations, d->resolution), "nz_thermal_erosion");
    }
};

class ErosionStageSubtractiveFlow : public PipelineStage {   // Geologic/Stage/ErosionStageSubtractiveFlow.cs:17-247 (commented out upstream)
public:
    int flowIterations = 5;                              // serialised upstream, read by nothing (:19-20, :226-228)
    float normMin = -0.1f, normMax = 0.1f, erosiveFactor = 0.1f;
    int erosiveIterations = 5;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_subtractive_flow_erosion(d->data, d->resolution, erosiveIterations, erosiveFactor, normMin, normMax),
              "nz_subtractive_flow_erosion");
    }
};

class ConstantStage : public PipelineStage {            // Filter/ConstantStage.cs:13-60
public:
    ConstantOperationType operation = ConstantOperationType::MULTIPLY;
    float value = 0.5f;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_constant(d->data, nz_slice_f32{nullptr, 0, 0}, (int)operation, value, d->resolution), "nz_constant");
    }
};

class ReduceStage : public PipelineStage {              // Filter/Reduce/ReduceStage.cs:21-68
    GeneratorData transformed;

public:
    ReductionType operation = ReductionType::SUBTRACT;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        ReduceData* d = CheckRequirements<ReduceData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_reduce(d->data, d->rightData, nz_slice_f32{nullptr, 0, 0}, (int)operation, d->resolution), "nz_reduce");
    }
    void TransformData(PipelineWorkItem& inputData) override {   // downstream stages see a GeneratorData (:52-61)
        ReduceData* d = static_cast<ReduceData*>(inputData.data);
        transformed.uuid = d->uuid; transformed.data = d->data; transformed.resolution = d->resolution;
        transformed.xpos = d->xpos; transformed.zpos = d->zpos;
        inputData.data = &transformed;
    }
};

class CurveStage : public PipelineStage {               // Filter/Curve/CurveStage.cs:13-73
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
        jobHandle = JobHandle::Chain(dependency);
        check(nz_curve(d->data, nz_slice_f32{nullptr, 0, 0}, nz_slice_f32{curve.data(), 4, (int32_t)curve.size()}, d->resolution), "nz_curve");
    }
};

class CropStage : public PipelineStage {                // Filter/Sample/CropStage.cs:13-19
public:
    bool center = false;   // false: the reference's behaviour (CropJob.Offset is never assigned: top-left corner)
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        DownsampleData* d = dynamic_cast<DownsampleData*>(requirements.data);
        if (!d) throw std::runtime_error("Unhandled stageio");
        jobHandle = JobHandle::Chain(dependency);
        check(nz_crop(d->inputData, d->inputResolution, d->data, d->resolution, center ? (d->inputResolution - d->resolution) / 2 : 0), "nz_crop");
    }
};

class FlowMapStage : public PipelineStage {
public:
    int iterations = 5;
    float normMin = -.1f, normMax = .1f;
    void Schedule(PipelineWorkItem& requirements, JobHandle dependency) override {
        GeneratorData* d = CheckRequirements<GeneratorData>(requirements);
        jobHandle = JobHandle::Chain(dependency);
        check(nz_flowmap(d->data, d->resolution, iterations, normMin, normMax), "nz_flowmap");
    }
};

class MeshTileStage : public PipelineStage {
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
        jobHandle = JobHandle::Chain(dependency);
        check(nz_heightmap_mesh((int)meshType, m.vertices.data(), m.indices.data(), R, d->inputResolution, d->marginPix,
                                d->tileHeight, d->tileSize, d->data), "nz_heightmap_mesh");
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
