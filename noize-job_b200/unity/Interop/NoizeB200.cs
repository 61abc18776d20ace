// NoizeB200.cs — C# host side of the B200 path: P/Invoke bindings of libnoize_b200.so plus GPU stages that
// subclass the reference's PipelineStage, so they drop into an existing PipelineDefinition unchanged.
//
// Lives next to Interop/MSWrapper.cs (asmdef xshazwar.noize.interop).  NOT compiled in this repository's
// image (no Unity / mono / dotnet there); the Python mirror (noize-job_b200/stages.py) exercises the same
// C ABI with the same structure and is what the tests run.  Keep the two files in step.
//
// Contract kept from the reference (Pipeline/Stage/PipelineStage.cs:41-57): Schedule() must not block and
// must set `jobHandle` to a handle that completes only when `d.data` (host memory) holds the result.
// Each GPU stage therefore schedules one IJob whose Execute() makes the blocking native call on a worker
// thread, exactly where the Burst job body used to run.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Jobs;
using UnityEngine;
using UnityEngine.Rendering;

using xshazwar.noize.pipeline;
using xshazwar.noize.filter;
using xshazwar.noize.filter.blur;
using xshazwar.noize.generate;
using xshazwar.noize.mesh;

namespace xshazwar.noize.interop.b200 {

    /// nz_slice_f32 of include/noize_b200.h == NativeSlice<float> (ptr, stride, length)
    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct NzSlice {
        public void* ptr;
        public int strideBytes;
        public int length;

        public static NzSlice From(NativeSlice<float> s) => new NzSlice {
            ptr = s.GetUnsafePtr(), strideBytes = s.Stride, length = s.Length
        };
        public static NzSlice Null => new NzSlice { ptr = null, strideBytes = 0, length = 0 };
    }

    public static unsafe class Native {
        const string LIB = "noize_b200";   // libnoize_b200.so in Assets/Plugins/x86_64

        [DllImport(LIB)] public static extern int nz_init(int* devices, int n);
        // multi-GPU: large grids (>= 4096 rows) handed to the stage calls below are split into row bands over the first
        // nBands devices of nz_init, inside the library; the stages do not change (include/noize_b200.h, "multi-GPU")
        [DllImport(LIB)] public static extern int nz_set_bands(int nBands);
        [DllImport(LIB)] public static extern int nz_shutdown();
        [DllImport(LIB)] public static extern IntPtr nz_last_error();
        [DllImport(LIB)] public static extern IntPtr nz_version();

        // FractalJobDelegate, Noise/Fractal/Fractal.cs:76-88
        [DllImport(LIB)] public static extern int nz_fractal(NzSlice dst, int resolution, int noiseType, float hurst,
            float startingAmplitude, float stepdown, float detuneRate, int octaves, int xpos, int zpos, int noiseSize);
        // SeperableKernelFilterDelegate x iterations, Filter/Kernel/KernelJob.cs:308-314 + KernelFilterStage.cs:31-43
        [DllImport(LIB)] public static extern int nz_kernel_filter(NzSlice src, NzSlice tmp, int filterType, int resolution, int iterations);
        // GaussFilterDelegate / SmoothFilterDelegate, Filter/Kernel/Blur/BlurJob.cs:23-30,46-52
        [DllImport(LIB)] public static extern int nz_gauss_filter(NzSlice src, NzSlice tmp, int width, int sigma, int resolution, int iterations);
        [DllImport(LIB)] public static extern int nz_smooth_filter(NzSlice src, NzSlice tmp, int width, int resolution, int iterations);
        // ErosionKernelJobDelegate x iterations, KernelJob.cs:350
        [DllImport(LIB)] public static extern int nz_min_erosion(NzSlice src, int resolution, int iterations);
        // FlowMapStage.ScheduleAll, Geologic/Stage/FlowMapStage.cs:124-195
        [DllImport(LIB)] public static extern int nz_flowmap(NzSlice height, int resolution, int iterations, float normMin, float normMax);
        // HeightMapMeshJobScheduleDelegate, Mesh/Job/HeightMapMeshJob.cs:55-65
        [DllImport(LIB)] public static extern int nz_heightmap_mesh(int meshType, void* vertices, uint* indices, int resolution,
            int inputResolution, int marginPix, float tileHeight, float tileSize, NzSlice heights);

        // ThermalErosionFilterDelegate, Filter/Kernel/Blur/ThermalErosionFilter.cs:138-146
        [DllImport(LIB)] public static extern int nz_thermal_erosion(NzSlice src, float talus, float incrementRatio,
            float meshHeightWidthRatio, int iterations, int resolution);
        // ErosionStageSubtractiveFlow.ScheduleAll, Geologic/Stage/ErosionStageSubtractiveFlow.cs:224-230 (cycle :138-222)
        [DllImport(LIB)] public static extern int nz_subtractive_flow_erosion(NzSlice height, int resolution, int erosiveIterations,
            float erosiveFactor, float normMin, float normMax);
        // ConstantJobScheduleDelegate, Filter/ConstantJob.cs:49-55
        [DllImport(LIB)] public static extern int nz_constant(NzSlice src, NzSlice tmp, int operation, float constantValue, int resolution);
        // ReductionJobScheduleDelegate, Filter/ReductionJob.cs:55-61
        [DllImport(LIB)] public static extern int nz_reduce(NzSlice left, NzSlice right, NzSlice tmp, int operation, int resolution);
        // CurveJobScheduleDelegate, Filter/Curve/CurveJob.cs:91-97
        [DllImport(LIB)] public static extern int nz_curve(NzSlice src, NzSlice tmp, NzSlice curve, int resolution);
        // CropJobDelegate, Filter/Sample/CropJob.cs:63-69 (offset 0 == the reference)
        [DllImport(LIB)] public static extern int nz_crop(NzSlice input, int inputResolution, NzSlice output, int outputResolution, int offset);
        // GetMapRangeJob / MapNormalizeValuesDelegate, Filter/NormalizeJob.cs:18-53,94-100
        [DllImport(LIB)] public static extern int nz_map_range(NzSlice map, float* res3, float limMin, float limMax);
        [DllImport(LIB)] public static extern int nz_normalize(NzSlice src, NzSlice tmp, float* args3, int resolution);

        // named device-resident buffers: the GPU side of PipelineStateManager (Pipeline/PipelineState/PipelineStateManager.cs:39-127)
        [DllImport(LIB)] public static extern int nz_context_write([MarshalAs(UnmanagedType.LPStr)] string name, NzSlice src);
        [DllImport(LIB)] public static extern int nz_context_read([MarshalAs(UnmanagedType.LPStr)] string name, NzSlice dst);
        [DllImport(LIB)] public static extern int nz_context_exists([MarshalAs(UnmanagedType.LPStr)] string name, int* length);
        [DllImport(LIB)] public static extern int nz_context_download([MarshalAs(UnmanagedType.LPStr)] string name, float* dst, int length);
        [DllImport(LIB)] public static extern int nz_context_upload([MarshalAs(UnmanagedType.LPStr)] string name, float* src, int length);
        [DllImport(LIB)] public static extern int nz_context_release([MarshalAs(UnmanagedType.LPStr)] string name);

        // process-wide residency scope: the stages of one chain run on different worker threads
        [DllImport(LIB)] public static extern long nz_scope_create();
        [DllImport(LIB)] public static extern int nz_scope_enter(long scope);
        [DllImport(LIB)] public static extern int nz_scope_leave();
        [DllImport(LIB)] public static extern int nz_scope_close(long scope);
        [DllImport(LIB)] public static extern int nz_pipeline_begin();
        [DllImport(LIB)] public static extern int nz_pipeline_end();
        [DllImport(LIB)] public static extern int nz_flush_to_host(void* hostPtr);
        [DllImport(LIB)] public static extern int nz_pin(void* hostPtr, UIntPtr bytes);
        [DllImport(LIB)] public static extern int nz_unpin(void* hostPtr);

        public static void Check(int rc, string what) {
            if (rc < 0) throw new Exception($"noize_b200 {what} failed ({rc}): {Marshal.PtrToStringAnsi(nz_last_error())}");
        }
    }

    /// One blocking native call, run where the Burst job body used to run.  Not [BurstCompile]: P/Invoke from a
    /// managed IJob is legal; the status code comes back through a NativeReference checked in OnStageComplete.
    public unsafe struct NativeCallJob : IJob {
        public enum Op { Fractal, KernelFilter, GaussFilter, SmoothFilter, MinErosion, FlowMap, Mesh, ThermalErosion, Constant, Reduce, Curve, Crop, SubtractiveFlow }
        public Op op;
        [NativeDisableContainerSafetyRestriction] public NativeSlice<float> data;
        [NativeDisableContainerSafetyRestriction] public NativeSlice<float> data2;   // right operand / curve samples / crop input
        [NativeDisableUnsafePtrRestriction] public void* vertices;
        [NativeDisableUnsafePtrRestriction] public uint* indices;
        public int resolution, inputResolution, marginPix, i0, i1, i2, xpos, zpos;
        public float f0, f1, f2, f3;
        [NativeDisableContainerSafetyRestriction] public NativeReference<int> status;
        public long scope;   // 0: stand-alone call (H2D + D2H around this stage); else a scope shared by the chain's jobs

        public void Execute() {
            NzSlice s = NzSlice.From(data);
            int rc = 0;
            if (scope != 0 && (rc = Native.nz_scope_enter(scope)) < 0) { status.Value = rc; return; }
            switch (op) {
                case Op.Fractal:      rc = Native.nz_fractal(s, resolution, i0, f0, f1, f2, f3, i1, xpos, zpos, i2); break;
                case Op.KernelFilter: rc = Native.nz_kernel_filter(s, NzSlice.Null, i0, resolution, i1); break;
                case Op.GaussFilter:  rc = Native.nz_gauss_filter(s, NzSlice.Null, i0, i2, resolution, i1); break;
                case Op.SmoothFilter: rc = Native.nz_smooth_filter(s, NzSlice.Null, i0, resolution, i1); break;
                case Op.MinErosion:   rc = Native.nz_min_erosion(s, resolution, i1); break;
                case Op.FlowMap:      rc = Native.nz_flowmap(s, resolution, i1, f0, f1); break;
                case Op.Mesh:         rc = Native.nz_heightmap_mesh(i0, vertices, indices, resolution, inputResolution, marginPix, f0, f1, s); break;
                case Op.ThermalErosion: rc = Native.nz_thermal_erosion(s, f0, f1, f2, i1, resolution); break;
                case Op.SubtractiveFlow: rc = Native.nz_subtractive_flow_erosion(s, resolution, i1, f2, f0, f1); break;
                case Op.Constant:     rc = Native.nz_constant(s, NzSlice.Null, i0, f0, resolution); break;
                case Op.Reduce:       rc = Native.nz_reduce(s, NzSlice.From(data2), NzSlice.Null, i0, resolution); break;
                case Op.Curve:        rc = Native.nz_curve(s, NzSlice.Null, NzSlice.From(data2), resolution); break;
                case Op.Crop:         rc = Native.nz_crop(NzSlice.From(data2), inputResolution, s, resolution, i0); break;
            }
            if (scope != 0) Native.nz_scope_leave();
            status.Value = rc;
        }
    }

    /// Closes a residency scope when the chain's last job has run: flushes the device mirrors to the host slices.
    /// Scheduled by the last GPU stage with its own job as dependency; ITS handle is what the stage hands on.
    public struct CloseScopeJob : IJob {
        public long scope;
        [NativeDisableContainerSafetyRestriction] public NativeReference<int> status;
        public void Execute() {
            int rc = Native.nz_scope_close(scope);
            if (status.Value >= 0) status.Value = rc;      // keep the first failure of the chain
        }
    }

    /// keepResident: brings one slice home (the handle's contract: host memory holds the result) but leaves the scope,
    /// and so the tile, in HBM for the next pipeline that works on the same uuid.
    public unsafe struct FlushScopeJob : IJob {
        public long scope;
        [NativeDisableContainerSafetyRestriction] public NativeSlice<float> data;
        [NativeDisableContainerSafetyRestriction] public NativeReference<int> status;
        public void Execute() {
            int rc = Native.nz_scope_enter(scope);
            if (rc >= 0) {
                rc = Native.nz_flush_to_host(data.GetUnsafePtr());
                Native.nz_scope_leave();
            }
            if (status.Value >= 0) status.Value = rc;
        }
    }

    /// uuid -> residency scope shared by the GPU stages that work on ONE work item (StageIO.uuid).  Main thread only:
    /// Schedule() and OnStageScheduled() always run there (Pipeline/Stage/PipelineStage.cs:44-57).
    public static class GpuResidency {
        static readonly Dictionary<string, long> scopes = new Dictionary<string, long>();

        public static long Enter(string uuid) {
            if (!scopes.TryGetValue(uuid, out long scope)) {
                scope = Native.nz_scope_create();
                if (scope < 0) Native.Check((int) scope, "nz_scope_create");
                scopes[uuid] = scope;
            }
            return scope;
        }
        /// the scope leaves the table now; it is closed by the job this returns (0: nothing open for uuid)
        public static long Detach(string uuid) {
            if (!scopes.TryGetValue(uuid, out long scope)) return 0;
            scopes.Remove(uuid);
            return scope;
        }
        public static long Peek(string uuid) => scopes.TryGetValue(uuid, out long scope) ? scope : 0;
        /// e.g. from a MonoBehaviour's OnDestroy: nothing may stay resident when the pipeline objects go away
        public static void CloseAll() {
            foreach (long scope in scopes.Values) Native.nz_scope_close(scope);
            scopes.Clear();
        }
    }

    /// Base of every stage that runs on the GPU.  Stage objects, not the pipeline, own residency:
    ///  * Schedule(): the first GPU stage of a work item creates the scope (GpuResidency.Enter), every stage passes it to
    ///    its NativeCallJob, which brackets the native call with nz_scope_enter / nz_scope_leave on its worker thread;
    ///  * OnStageScheduled(): when the next receiver is NOT a GPU stage — a Burst stage, or BasePipeline's
    ///    OnPipelineFullyScheduled (Pipeline/Executable/Pipeline.cs:122-151) — the scope's close (or, with keepResident,
    ///    a flush of this stage's slice) is chained behind this stage's job and ITS handle is handed on, so the handle
    ///    still completes only when d.data (host memory) holds the result.
    /// Chained GPU stages therefore pay one H2D (none after a generator) and one D2H per work item, not per stage.
    public abstract class GpuStage : PipelineStage {
        [Tooltip("Last GPU stage of a pipeline: bring the result home but keep the tile in HBM for the next pipeline on the same uuid (e.g. generator pipeline -> mesh pipeline)")]
        public bool keepResident = false;
        protected NativeReference<int> status;
        protected void EnsureStatus() {
            if (!status.IsCreated) status = new NativeReference<int>(Allocator.Persistent);
            status.Value = 0;
        }
        protected long EnterScope(StageIO d) {
            EnsureStatus();
            return GpuResidency.Enter(d.uuid);
        }
        bool NextIsGpuStage() {
            if (OnStageScheduledAction == null) return false;
            foreach (Delegate next in OnStageScheduledAction.GetInvocationList())
                if (next.Target is GpuStage) return true;
            return false;
        }
        public override void OnStageScheduled(PipelineWorkItem requirements, JobHandle dependency) {
            if (!NextIsGpuStage()) {
                string uuid = requirements.data.uuid;
                if (keepResident) {
                    long scope = GpuResidency.Peek(uuid);
                    if (scope != 0) jobHandle = new FlushScopeJob { scope = scope, data = requirements.data.data, status = status }.Schedule(jobHandle);
                } else {
                    long scope = GpuResidency.Detach(uuid);
                    if (scope != 0) jobHandle = new CloseScopeJob { scope = scope, status = status }.Schedule(jobHandle);
                }
            }
            OnStageScheduledAction?.Invoke(requirements, jobHandle);
        }
        public override void OnStageComplete() {
            if (status.IsCreated && status.Value < 0) Native.Check(status.Value, GetType().Name);
        }
        public override void OnDestroy() { if (status.IsCreated) status.Dispose(); }
    }

    [CreateAssetMenu(fileName = "GpuNoiseGenerator", menuName = "Noize/B200/NoiseSource", order = 1)]
    public class GpuNoiseStage : GpuStage {
        public NoiseStage.FractalNoise noiseType;
        [Range(0f, 2f)] public float hurst = 0f;
        [Range(.01f, 5f)] public float startingAmplitude = 1f;
        [Range(1, 24)] public int octaves = 1;
        [Range(1.8f, 2.2f)] public float stepdown = 2f;
        [Range(-.05f, .05f)] public float detuneRate = 0f;
        [Range(5, 32000)] public int noiseSize = 1000;

        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.Fractal, data = d.data, resolution = d.resolution, i0 = (int) noiseType, f0 = hurst,
                f1 = startingAmplitude, f2 = stepdown, f3 = detuneRate, i1 = octaves, xpos = d.xpos, zpos = d.zpos,
                i2 = noiseSize, status = status
            }.Schedule(dependency);
        }
    }

    [CreateAssetMenu(fileName = "GpuKernelFilter", menuName = "Noize/B200/KernelFilter", order = 2)]
    public class GpuKernelFilterStage : GpuStage {
        public KernelFilterType filter;
        [Range(1, 32)] public int iterations = 1;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            // the reference chains `iterations` jobs; the GPU stage issues ONE fused call
            jobHandle = new NativeCallJob {
                scope = scope, op = NativeCallJob.Op.KernelFilter, data = d.data, resolution = d.resolution, i0 = (int) filter,
                i1 = iterations, status = status
            }.Schedule(dependency);
        }
    }

    // StageGaussianBlur, Filter/Kernel/Blur/StageGaussianBlur.cs:12-56 (same serialized fields; limitWidth is applied natively,
    // BlurKernels.cs:30-36).  The reference chains `iterations` GaussFilter jobs (:33-46); this is one fused call.
    [CreateAssetMenu(fileName = "GpuStageGaussianBlur", menuName = "Noize/B200/Blur/GaussianBlurFilter", order = 2)]
    public class GpuStageGaussianBlur : GpuStage {
        [Range(1, 32)] public int iterations = 1;
        public GaussSigma sigma;
        [Range(3, 25)] public int width = 3;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope, op = NativeCallJob.Op.GaussFilter, data = d.data, resolution = d.resolution, i0 = width,
                i2 = (int) sigma, i1 = iterations, status = status
            }.Schedule(dependency);
        }
    }

    // StageSmoothBlur, Filter/Kernel/Blur/StageSmoothBlur.cs:12-54 (box blur of odd width, 1/width taps)
    [CreateAssetMenu(fileName = "GpuStageSmoothBlur", menuName = "Noize/B200/Blur/SmoothBlurFilter", order = 2)]
    public class GpuStageSmoothBlur : GpuStage {
        [Range(1, 32)] public int iterations = 1;
        [Range(3, 25)] public int width = 1;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope, op = NativeCallJob.Op.SmoothFilter, data = d.data, resolution = d.resolution, i0 = width,
                i1 = iterations, status = status
            }.Schedule(dependency);
        }
    }

    [CreateAssetMenu(fileName = "GpuErosionFilter", menuName = "Noize/B200/ValueErosion", order = 2)]
    public class GpuErosionFilterStage : GpuStage {
        [Range(1, 32)] public int iterations = 5;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.MinErosion, data = d.data, resolution = d.resolution, i1 = iterations, status = status
            }.Schedule(dependency);
        }
    }

    // StageThermalErosion, Filter/Kernel/Blur/StageThermalErosion.cs:13-29 (same serialized fields)
    [CreateAssetMenu(fileName = "GpuStageThermalErosion", menuName = "Noize/B200/ThermalErosion", order = 2)]
    public class GpuStageThermalErosion : GpuStage {
        [Range(1, 32)] public int iterations = 1;
        [Range(1, 90)] public int talus = 45;
        public float increment = 0.5f;
        public float meshHeightWidthRatio = 0.75f;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.ThermalErosion, data = d.data, resolution = d.resolution, i1 = iterations,
                f0 = (float) talus, f1 = increment, f2 = meshHeightWidthRatio, status = status
            }.Schedule(dependency);
        }
    }

    // ErosionStageSubtractiveFlow, Geologic/Stage/ErosionStageSubtractiveFlow.cs:17-247 (commented out upstream; same fields)
    [CreateAssetMenu(fileName = "GpuErosionSubtractiveFlowStage", menuName = "Noize/B200/SubtractiveFlow", order = 2)]
    public class GpuErosionStageSubtractiveFlow : GpuStage {
        [Range(1, 32)] public int flowIterations = 5;   // unused upstream as well: cycle n runs n + 1 iterations
        public float normMin = -.1f;
        public float normMax = .1f;
        public float erosiveFactor = .1f;
        public int erosiveIterations = 5;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.SubtractiveFlow, data = d.data, resolution = d.resolution, i1 = erosiveIterations,
                f0 = normMin, f1 = normMax, f2 = erosiveFactor, status = status
            }.Schedule(dependency);
        }
    }

    // ConstantStage, Filter/ConstantStage.cs:13-60
    [CreateAssetMenu(fileName = "GpuConstant", menuName = "Noize/B200/Constant", order = 2)]
    public class GpuConstantStage : GpuStage {
        public ConstantStage.ConstantOperationType operation;
        [Range(0, 1)] public float value = 0.5f;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.Constant, data = d.data, resolution = d.resolution, i0 = (int) operation, f0 = value, status = status
            }.Schedule(dependency);
        }
    }

    // ReduceStage, Filter/Reduce/ReduceStage.cs:21-68 (TransformData hands a GeneratorData downstream, :52-61)
    [CreateAssetMenu(fileName = "GpuReduceStage", menuName = "Noize/B200/ReduceFilter", order = 2)]
    public class GpuReduceStage : GpuStage {
        public ReductionType operation;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<ReduceData>(requirements);
            ReduceData d = (ReduceData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.Reduce, data = d.data, data2 = d.rightData, resolution = d.resolution, i0 = (int) operation, status = status
            }.Schedule(dependency);
        }
        public override void TransformData(PipelineWorkItem inputData) {
            ReduceData d = (ReduceData) inputData.data;
            inputData.data = new GeneratorData { uuid = d.uuid, resolution = d.resolution, data = d.data, xpos = d.xpos, zpos = d.zpos };
        }
    }

    // CurveStage, Filter/Curve/CurveStage.cs:13-73: the AnimationCurve is discretised exactly as ExtractCurve does (:27-35)
    [CreateAssetMenu(fileName = "GpuCurveStage", menuName = "Noize/B200/CurveFilter", order = 2)]
    public class GpuCurveStage : GpuStage {
        public AnimationCurve unityCurve;
        public int samples = 256;
        NativeArray<float> curve;
        public override void ResizeNativeContainers(int size) {
            if (curve.IsCreated) curve.Dispose();
            curve = new NativeArray<float>(samples, Allocator.Persistent, NativeArrayOptions.UninitializedMemory);
            for (int i = 0; i < samples; i++) curve[i] = unityCurve.Evaluate((float) i / samples);
        }
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.Curve, data = d.data, data2 = new NativeSlice<float>(curve), resolution = d.resolution, status = status
            }.Schedule(dependency);
        }
        public override void OnDestroy() { if (curve.IsCreated) curve.Dispose(); base.OnDestroy(); }
    }

    // CropStage, Filter/Sample/CropStage.cs:13-19 (offset 0 reproduces the reference, whose CropJob.Offset is never set)
    [CreateAssetMenu(fileName = "GpuCropStage", menuName = "Noize/B200/CenterCropResolution", order = 2)]
    public class GpuCropStage : GpuStage {
        public bool center = false;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            DownsampleData d = (DownsampleData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.Crop, data = d.data, data2 = d.inputData, resolution = d.resolution,
                inputResolution = d.inputResolution, i0 = center ? (d.inputResolution - d.resolution) / 2 : 0, status = status
            }.Schedule(dependency);
        }
    }

    /// Write/ReadGeneratorContextStage on the GPU: the tile is parked in (or fetched from) a NAMED buffer in HBM that outlives
    /// residency scopes — one device-to-device copy where the reference runs a FlushWriteSlice memcpy job between two host
    /// arrays (PipelineState/Stage/WriteGeneratorContextStage.cs:36-52, ReadGeneratorContextStage.cs:39-51).  The job is
    /// managed (a string cannot cross into Burst); the name is pinned in a GCHandle for the job's lifetime.
    public unsafe struct ContextCopyJob : IJob {
        public bool write;
        public GCHandle name;
        [NativeDisableContainerSafetyRestriction] public NativeSlice<float> data;
        [NativeDisableContainerSafetyRestriction] public NativeReference<int> status;
        public long scope;
        public void Execute() {
            string n = (string) name.Target;
            int rc = scope != 0 ? Native.nz_scope_enter(scope) : 0;
            if (rc >= 0) {
                rc = write ? Native.nz_context_write(n, NzSlice.From(data)) : Native.nz_context_read(n, NzSlice.From(data));
                if (scope != 0) Native.nz_scope_leave();
            }
            name.Free();
            status.Value = rc;
        }
    }

    public abstract class GpuContextStage : GpuStage {
        public string contextAlias;
        protected string BufferName(GeneratorData d) => $"{d.xpos}_{d.zpos}__{d.resolution}__{contextAlias}";   // getBufferName, :21-23
        protected static bool Exists(string name) { int n = -1; Native.nz_context_exists(name, &n); return n >= 0; }
        protected void ScheduleCopy(PipelineWorkItem requirements, JobHandle dependency, bool write) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new ContextCopyJob {
                write = write, name = GCHandle.Alloc(BufferName(d)), data = d.data, status = status, scope = scope
            }.Schedule(dependency);
        }
    }

    [CreateAssetMenu(fileName = "GpuWriteGeneratorContextStage", menuName = "Noize/B200/State/WriteContext", order = 2)]
    public unsafe class GpuWriteGeneratorContextStage : GpuContextStage {
        JobHandle pending;
        public override bool IsSchedulable(PipelineWorkItem job) => pending.IsCompleted;          // the LockJob of the reference (:43-52)
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            ScheduleCopy(requirements, dependency, true);
            pending = jobHandle;
        }
    }

    [CreateAssetMenu(fileName = "GpuReadGeneratorContextStage", menuName = "Noize/B200/State/ReadContext", order = 2)]
    public unsafe class GpuReadGeneratorContextStage : GpuContextStage {
        public override bool IsSchedulable(PipelineWorkItem job) => Exists(BufferName((GeneratorData) job.data));   // :24-37
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            ScheduleCopy(requirements, dependency, false);
        }
    }

    [CreateAssetMenu(fileName = "GpuFlowMapStage", menuName = "Noize/B200/FlowMap", order = 2)]
    public class GpuFlowMapStage : GpuStage {
        [Range(1, 128)] public int iterations = 5;
        public float normMin = -.1f;
        public float normMax = .1f;
        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            CheckRequirements<GeneratorData>(requirements);
            GeneratorData d = (GeneratorData) requirements.data;
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.FlowMap, data = d.data, resolution = d.resolution, i1 = iterations, f0 = normMin,
                f1 = normMax, status = status
            }.Schedule(dependency);
        }
    }

    [CreateAssetMenu(fileName = "GpuMeshTileStage", menuName = "Noize/B200/MeshTile", order = 2)]
    public unsafe class GpuMeshTileStage : GpuStage {
        public MeshType meshType = MeshType.SquareGridHeightMap;
        Mesh currentMesh;
        Mesh.MeshDataArray meshDataArray;

        public override void Schedule(PipelineWorkItem requirements, JobHandle dependency) {
            MeshStageData d = (MeshStageData) requirements.data;
            currentMesh = d.mesh;
            meshDataArray = Mesh.AllocateWritableMeshData(1);
            Mesh.MeshData md = meshDataArray[0];
            int R = d.resolution, vcount = (R + 1) * (R + 1), icount = 6 * R * R;
            // same declaration as PositionStream32.Setup, Mesh/Streams/PositionStream.cs:90-123
            var desc = new NativeArray<VertexAttributeDescriptor>(4, Allocator.Temp, NativeArrayOptions.UninitializedMemory);
            desc[0] = new VertexAttributeDescriptor(dimension: 3);
            desc[1] = new VertexAttributeDescriptor(VertexAttribute.Normal, dimension: 3);
            desc[2] = new VertexAttributeDescriptor(VertexAttribute.Tangent, dimension: 4);
            desc[3] = new VertexAttributeDescriptor(VertexAttribute.TexCoord0, dimension: 2);
            md.SetVertexBufferParams(vcount, desc);
            desc.Dispose();
            md.SetIndexBufferParams(icount, IndexFormat.UInt32);
            var bounds = new Bounds(new Vector3(0.5f * d.tileSize, 0.5f * d.tileHeight, 0.5f * d.tileSize),
                                    new Vector3(d.tileSize, d.tileHeight, d.tileSize));
            currentMesh.bounds = bounds;
            md.subMeshCount = 1;
            md.SetSubMesh(0, new SubMeshDescriptor(0, icount) { bounds = bounds, vertexCount = vcount },
                          MeshUpdateFlags.DontRecalculateBounds | MeshUpdateFlags.DontValidateIndices);
            long scope = EnterScope(d);
            jobHandle = new NativeCallJob {
                scope = scope,
                op = NativeCallJob.Op.Mesh, data = d.data, vertices = md.GetVertexData<byte>().GetUnsafePtr(),
                indices = (uint*) md.GetIndexData<uint>().GetUnsafePtr(), resolution = R, inputResolution = d.inputResolution,
                marginPix = d.marginPix, i0 = (int) meshType, f0 = d.tileHeight, f1 = d.tileSize, status = status
            }.Schedule(dependency);
        }

        public override void OnStageComplete() {
            base.OnStageComplete();
            Mesh.ApplyAndDisposeWritableMeshData(meshDataArray, currentMesh,
                MeshUpdateFlags.DontNotifyMeshUsers | MeshUpdateFlags.DontValidateIndices | MeshUpdateFlags.DontRecalculateBounds);
        }
    }
}
