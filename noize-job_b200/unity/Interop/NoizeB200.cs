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
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Jobs;
using UnityEngine;
using UnityEngine.Rendering;

using xshazwar.noize.pipeline;
using xshazwar.noize.filter;
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
            if (scope != 0) Native.nz_scope_enter(scope);
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
    /// Schedule it with the last stage's JobHandle as dependency and hand ITS handle to the pipeline.
    public struct CloseScopeJob : IJob {
        public long scope;
        [NativeDisableContainerSafetyRestriction] public NativeReference<int> status;
        public void Execute() { status.Value = Native.nz_scope_close(scope); }
    }

    public abstract class GpuStage : PipelineStage {
        protected NativeReference<int> status;
        protected void EnsureStatus() {
            if (!status.IsCreated) status = new NativeReference<int>(Allocator.Persistent);
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            // the reference chains `iterations` jobs; the GPU stage issues ONE fused call
            jobHandle = new NativeCallJob {
                op = NativeCallJob.Op.KernelFilter, data = d.data, resolution = d.resolution, i0 = (int) filter,
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
                op = NativeCallJob.Op.Crop, data = d.data, data2 = d.inputData, resolution = d.resolution,
                inputResolution = d.inputResolution, i0 = center ? (d.inputResolution - d.resolution) / 2 : 0, status = status
            }.Schedule(dependency);
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
            EnsureStatus();
            jobHandle = new NativeCallJob {
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
