"""Host-side mirror of the reference's Pipeline stage API, bound to the C ABI.

Same names, fields, defaults and error behaviour as the C# classes, so a reference user can read this
file side by side with the originals (paths relative to xshazwar/noize-job):

  PipelineStage / IStage            Pipeline/Stage/PipelineStage.cs:10-62, Pipeline/Interface.cs:7-45
  StageIO, GeneratorData, ...       Pipeline/Stage/StageIO.cs:8-11, Pipeline/Stage/StageIOTypes/*.cs
  PipelineWorkItem                  Pipeline/Stage/PipelineDefinition.cs:18-25
  NoiseStage                        Noise/NoiseStage.cs:13-61
  KernelFilterStage                 Filter/KernelFilterStage.cs:13-51
  StageGaussianBlur/StageSmoothBlur Filter/Kernel/Blur/StageGaussianBlur.cs, StageSmoothBlur.cs
  FlowMapStage                      Geologic/Stage/FlowMapStage.cs:16-221
  MeshTileStage                     Mesh/Stage/MeshTileStage.cs:28-62
  BasePipeline (scheduling chain)   Pipeline/Executable/Pipeline.cs:104-181
  StageThermalErosion               Filter/Kernel/Blur/StageThermalErosion.cs:13-29
  ErosionStageSubtractiveFlow       Geologic/Stage/ErosionStageSubtractiveFlow.cs:17-247 (commented out upstream)
  ConstantStage                     Filter/ConstantStage.cs:13-60
  ReduceStage, ReduceData           Filter/Reduce/ReduceStage.cs:12-68, Pipeline/Stage/StageIOTypes/ReduceData.cs
  CurveStage                        Filter/Curve/CurveStage.cs:13-73
  CropStage, DownsampleData         Filter/Sample/CropStage.cs:13-19, Pipeline/Stage/StageIOTypes/DownsampleData.cs
  ErosionFilterStage                NEW: wraps ErosionKernelJob (Filter/Kernel/KernelJob.cs:317-350), which no
                                    stage binds in the current reference tree ("Value Erosion", README.md:18)

The C# production shim (unity/Interop/NoizeB200.cs) has the same structure with [DllImport] stubs.
A `NativeSlice<float>` is a 1-D numpy float32 view here (it may be strided).  `JobHandle.Complete()`
returns when the host buffers hold the results, as the reference's handle does.
"""
import threading
from collections import deque
from enum import IntEnum

import numpy as np

from . import host as _h


# ---- enums (numeric order == the C# enums == include/noize_b200.h) --------------------------------
class FractalNoise(IntEnum):           # Noise/NoiseStage.cs:15-24
    Sin = 0
    Perlin = 1
    PeriodicPerlin = 2
    Simplex = 3
    RotatedSimplex = 4
    Cellular = 5
    DomainRotatedPerlin = 6
    DomainRotatedSimplex = 7


class KernelFilterType(IntEnum):       # Filter/Kernel/KernelJob.cs:79-94
    Gauss9_S1 = 0
    Gauss7_S1 = 1
    Gauss5_S1 = 2
    Gauss3_S1 = 3
    Gauss9_S2 = 4
    Gauss7_S2 = 5
    Gauss5_S2 = 6
    Gauss3_S2 = 7
    Smooth3 = 8
    Sobel3Horizontal = 9
    Sobel3Vertical = 10
    Sobel3_2D = 11
    Prewitt3Horizontal = 12
    Prewitt3Vertical = 13


class GaussSigma(IntEnum):             # Filter/Kernel/Blur/BlurKernels.cs:8-25
    s0d50 = 0; s1d00 = 1; s1d50 = 2; s2d00 = 3; s2d50 = 4; s3d00 = 5; s3d50 = 6; s4d00 = 7
    s4d50 = 8; s5d00 = 9; s5d50 = 10; s6d00 = 11; s6d50 = 12; s7d00 = 13; s7d50 = 14; s8d00 = 15


class MeshType(IntEnum):               # Mesh/Stage/MeshTileStage.cs:22-25
    SquareGridHeightMap = 0
    OvershootSquareGridHeightMap = 1


class ConstantOperationType(IntEnum):  # Filter/ConstantStage.cs:15-18
    MULTIPLY = 0
    BINARIZE = 1


class ReductionType(IntEnum):          # Filter/Reduce/ReduceStage.cs:12-18
    SUBTRACT = 0
    MULTIPLY = 1
    ROOTSUMSQUARES = 2
    MAX = 3
    MIN = 4


# ---- job handle + residency (unity/Interop/NoizeB200.cs: GpuResidency, NativeCallJob, CloseScopeJob, FlushScopeJob) ----
class JobHandle:
    """Completion token of scheduled work.  The native calls of this mirror run inline where Unity would run an IJob on a
    worker thread, so a handle is complete when it is returned; `Complete()` is kept for the reference's call sites."""

    def __init__(self):
        self.IsCompleted = True

    def Complete(self):
        self.IsCompleted = True


class GpuResidency:
    """uuid -> residency scope (nz_scope_*) shared by the GPU stages that work on ONE work item (StageIO.uuid).

    The FIRST GPU stage of a work item creates the scope, every GPU stage brackets its native call with
    nz_scope_enter / nz_scope_leave (the call may run on any worker thread), and the LAST GPU stage before a non-GPU consumer
    (a Burst stage, or the pipeline's fully-scheduled hook, Pipeline/Executable/Pipeline.cs:122-128) chains the scope's close
    (one D2H per dirty slice) in front of the handle it hands on — or, with `keepResident`, only a flush, so that the next
    pipeline working on the same uuid (the reference runs the generator pipeline and the mesh pipeline back to back,
    Scripts/MeshTileGenerator.cs:94-138) finds the tile still in HBM."""
    _scopes = {}
    _lock = threading.Lock()

    @classmethod
    def Enter(cls, uuid):
        with cls._lock:
            scope = cls._scopes.get(uuid)
            if scope is None:
                scope = cls._scopes[uuid] = _h.scope_create()
            return scope

    @classmethod
    def Close(cls, uuid):
        with cls._lock:
            scope = cls._scopes.pop(uuid, None)
        if scope is not None:
            _h.scope_close(scope)

    @classmethod
    def Flush(cls, uuid, slices):
        with cls._lock:
            scope = cls._scopes.get(uuid)
        if scope is not None:
            with _h.in_scope(scope):
                for a in slices:
                    if a is not None:
                        _h.flush_to_host(a)

    @classmethod
    def IsOpen(cls, uuid):
        with cls._lock:
            return uuid in cls._scopes

    @classmethod
    def CloseAll(cls):
        with cls._lock:
            scopes, cls._scopes = list(cls._scopes.values()), {}
        for scope in scopes:
            _h.scope_close(scope)


# ---- stage IO ---------------------------------------------------------------------------------------
class StageIO:
    def __init__(self, uuid="", data=None):
        self.uuid = uuid
        self.data = data          # NativeSlice<float>: 1-D float32 numpy view


class GeneratorData(StageIO):
    def __init__(self, uuid="", data=None, resolution=512, xpos=0, zpos=0):
        super().__init__(uuid, data)
        self.resolution, self.xpos, self.zpos = resolution, xpos, zpos


class ReduceData(StageIO):             # Pipeline/Stage/StageIOTypes/ReduceData.cs
    def __init__(self, uuid="", data=None, rightData=None, resolution=512, xpos=0, zpos=0):
        super().__init__(uuid, data)
        self.rightData, self.resolution, self.xpos, self.zpos = rightData, resolution, xpos, zpos


class DownsampleData(StageIO):         # Pipeline/Stage/StageIOTypes/DownsampleData.cs
    def __init__(self, uuid="", data=None, resolution=512, inputResolution=512, inputData=None):
        super().__init__(uuid, data)
        self.resolution, self.inputResolution, self.inputData = resolution, inputResolution, inputData


class Mesh:
    """Stand-in for UnityEngine.Mesh + Mesh.MeshData: the two buffers PositionStream32.Setup declares
    (Mesh/Streams/PositionStream.cs:90-123) and the bounds HeightMapMeshJob assigns."""

    def __init__(self):
        self.vertices = None      # ((R+1)^2, 12) float32: pos3, normal3, tangent4, uv2
        self.indices = None       # 6*R^2 uint32
        self.bounds = None        # (center xyz, size xyz)


class MeshStageData(StageIO):
    def __init__(self, uuid="", data=None, resolution=512, inputResolution=512, marginPix=5, tileSize=512.0,
                 tileHeight=512.0, xpos=0, zpos=0, mesh=None):
        super().__init__(uuid, data)
        self.resolution, self.inputResolution, self.marginPix = resolution, inputResolution, marginPix
        self.tileSize, self.tileHeight, self.xpos, self.zpos, self.mesh = tileSize, tileHeight, xpos, zpos, mesh


class PipelineWorkItem:
    def __init__(self, data, completeAction=None, scheduledAction=None, dependency=None, stageManager=None):
        self.data, self.completeAction, self.scheduledAction = data, completeAction, scheduledAction
        self.dependency, self.stageManager = dependency, stageManager


# ---- stage base -------------------------------------------------------------------------------------
class PipelineStage:
    def __init__(self):
        self.jobHandle = None
        self.OnStageScheduledAction = None
        self.arraysInitialized = False
        self.dataLength = 0

    def ResizeNativeContainers(self, size):
        pass

    def IsSchedulable(self, job):
        return True

    def CheckRequirements(self, T, requirements):
        if isinstance(requirements.data, T):
            n = requirements.data.data.size
            if n != self.dataLength:
                self.dataLength = n
                self.ResizeNativeContainers(n)
        else:
            raise Exception(f"Unhandled stageio {type(requirements.data).__name__}")

    def Schedule(self, requirements, dependency):
        pass

    def ReceiveHandledInput(self, requirements, dependency):
        self.Schedule(requirements, dependency)
        self.TransformData(requirements)
        self.OnStageScheduled(requirements, self.jobHandle)

    def Destroy(self):
        self.OnDestroy()

    def TransformData(self, data):
        pass

    def OnStageScheduled(self, requirements, dependency):
        if self.OnStageScheduledAction is not None:
            self.OnStageScheduledAction(requirements, self.jobHandle)

    def OnStageComplete(self):
        pass

    def OnDestroy(self):
        pass


class GpuStage(PipelineStage):
    """Base of every stage that runs on the GPU (C#: abstract class GpuStage : PipelineStage).  Stage objects, not the
    pipeline, own the residency scope of the work item they are handed."""
    keepResident = False      # last GPU stage of a pipeline: flush the result to the host but leave the tile in HBM

    def _native(self, requirements, call):
        """NativeCallJob.Execute: enter the work item's scope, make the one blocking native call, leave.  A caller that
        already holds a scope on this thread (nz.host.pipeline()) keeps control of residency."""
        uuid = requirements.data.uuid
        if _h.thread_in_scope():
            call()
        else:
            scope = GpuResidency.Enter(uuid)
            try:
                with _h.in_scope(scope):
                    call()
            except Exception:
                GpuResidency.Close(uuid)      # CloseScopeJob still runs when a stage failed: nothing stays behind
                raise
        self.jobHandle = JobHandle()

    def _next_is_gpu_stage(self):
        nxt = getattr(self.OnStageScheduledAction, "__self__", None)
        return isinstance(nxt, GpuStage)

    def _slices(self, requirements):
        d = requirements.data
        return [getattr(d, "data", None)]

    def OnStageScheduled(self, requirements, dependency):
        # hand-over to something that reads HOST memory: the scope's close (or flush) goes in front of the handle
        if not self._next_is_gpu_stage() and not _h.thread_in_scope():
            uuid = requirements.data.uuid
            if self.keepResident:
                GpuResidency.Flush(uuid, self._slices(requirements))
            else:
                GpuResidency.Close(uuid)
        super().OnStageScheduled(requirements, dependency)


class NoiseStage(GpuStage):
    def __init__(self, noiseType=FractalNoise.Sin, hurst=0.0, startingAmplitude=1.0, octaves=1, stepdown=2.0,
                 detuneRate=0.0, noiseSize=1000):
        super().__init__()
        self.noiseType, self.hurst, self.startingAmplitude = noiseType, hurst, startingAmplitude
        self.octaves, self.stepdown, self.detuneRate, self.noiseSize = octaves, stepdown, detuneRate, noiseSize

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.fractal(d.data, d.resolution, self.noiseType, self.hurst, self.startingAmplitude,
                                                      self.stepdown, self.detuneRate, self.octaves, d.xpos, d.zpos, self.noiseSize))


class KernelFilterStage(GpuStage):
    def __init__(self, filter=KernelFilterType.Gauss9_S1, iterations=1):
        super().__init__()
        self.filter, self.iterations = filter, iterations

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        # the reference schedules `iterations` dependent jobs; the GPU stage issues ONE fused call
        self._native(requirements, lambda: _h.kernel_filter(d.data, None, self.filter, d.resolution, self.iterations))


class StageGaussianBlur(GpuStage):
    def __init__(self, iterations=1, sigma=GaussSigma.s0d50, width=3):
        super().__init__()
        self.iterations, self.sigma, self.width = iterations, sigma, width

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.gauss_filter(d.data, None, self.width, self.sigma, d.resolution, self.iterations))


class StageSmoothBlur(GpuStage):
    def __init__(self, iterations=1, width=1):
        super().__init__()
        self.iterations, self.width = iterations, width

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.smooth_filter(d.data, None, self.width, d.resolution, self.iterations))


class ErosionFilterStage(GpuStage):
    def __init__(self, iterations=5):
        super().__init__()
        self.iterations = iterations

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.min_erosion(d.data, d.resolution, self.iterations))


class StageThermalErosion(GpuStage):
    def __init__(self, iterations=1, talus=45, increment=0.5, meshHeightWidthRatio=0.75):
        super().__init__()
        self.iterations, self.talus, self.increment, self.meshHeightWidthRatio = iterations, talus, increment, meshHeightWidthRatio

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.thermal_erosion(d.data, float(self.talus), self.increment, self.meshHeightWidthRatio, self.iterations, d.resolution))


class ErosionStageSubtractiveFlow(GpuStage):
    """Fields and defaults of ErosionStageSubtractiveFlow.cs:19-27.  `flowIterations` is kept as a field because the
    reference serialises it, but nothing reads it there either: cycle n runs n + 1 flow iterations (:226-228)."""

    def __init__(self, flowIterations=5, normMin=-0.1, normMax=0.1, erosiveFactor=0.1, erosiveIterations=5):
        super().__init__()
        self.flowIterations, self.normMin, self.normMax = flowIterations, normMin, normMax
        self.erosiveFactor, self.erosiveIterations = erosiveFactor, erosiveIterations

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.subtractive_flow_erosion(d.data, d.resolution, self.erosiveIterations, self.erosiveFactor, self.normMin, self.normMax))


class ConstantStage(GpuStage):
    def __init__(self, operation=ConstantOperationType.MULTIPLY, value=0.5):
        super().__init__()
        self.operation, self.value = operation, value

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.constant(d.data, None, self.operation, self.value, d.resolution))


class ReduceStage(GpuStage):
    def __init__(self, operation=ReductionType.SUBTRACT):
        super().__init__()
        self.operation = operation

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(ReduceData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.reduce(d.data, d.rightData, None, self.operation, d.resolution))

    def _slices(self, requirements):
        return [requirements.data.data]

    def TransformData(self, inputData):    # ReduceStage.cs:52-61: downstream stages see a GeneratorData
        d = inputData.data
        inputData.data = GeneratorData(d.uuid, d.data, d.resolution, d.xpos, d.zpos)


class CurveStage(GpuStage):
    """`unityCurve` is any callable t -> value on [0,1] (AnimationCurve.Evaluate); it is discretised exactly as
    CurveStage.ExtractCurve does: curve[i] = Evaluate((float) i / samples), Filter/Curve/CurveStage.cs:27-35."""

    def __init__(self, unityCurve=None, samples=256):
        super().__init__()
        self.unityCurve, self.samples = unityCurve if unityCurve is not None else (lambda t: t), samples
        self.curve = None

    def ExtractCurve(self):
        s = np.float32(self.samples)
        self.curve = np.array([self.unityCurve(float(np.float32(i) / s)) for i in range(self.samples)], np.float32)

    def ResizeNativeContainers(self, size):
        self.ExtractCurve()

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.curve(d.data, None, self.curve, d.resolution))


class CropStage(GpuStage):
    """The reference never assigns CropJob.Offset, so its "CenterCropResolution" stage copies the top-left corner
    (Filter/Sample/CropJob.cs:25,36-43).  offset=None reproduces that; offset='center' is the evident intent."""

    def __init__(self, offset=None):
        super().__init__()
        self.offset = offset

    def Schedule(self, requirements, dependency):
        d = requirements.data
        if not isinstance(d, DownsampleData):
            raise Exception(f"Unhandled stageio {type(d).__name__}")
        off = 0 if self.offset is None else ((d.inputResolution - d.resolution) // 2 if self.offset == "center" else int(self.offset))
        self._native(requirements, lambda: _h.crop(d.inputData, d.inputResolution, d.data, d.resolution, off))


class FlowMapStage(GpuStage):
    def __init__(self, iterations=5, normMin=-0.1, normMax=0.1):
        super().__init__()
        self.iterations, self.normMin, self.normMax = iterations, normMin, normMax

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        d = requirements.data
        self._native(requirements, lambda: _h.flowmap(d.data, d.resolution, self.iterations, self.normMin, self.normMax))


class MeshTileStage(GpuStage):
    def __init__(self, meshType=MeshType.SquareGridHeightMap):
        super().__init__()
        self.meshType = meshType
        self.currentMesh = None

    def Schedule(self, requirements, dependency):
        d = requirements.data
        if not isinstance(d, MeshStageData):
            raise Exception(f"Unhandled stageio {type(d).__name__}")
        self.currentMesh = d.mesh if d.mesh is not None else Mesh()
        d.mesh = self.currentMesh
        R = d.resolution
        # Mesh.AllocateWritableMeshData(1) + PositionStream32.Setup; buffers the caller already attached to
        # the Mesh (e.g. pinned memory) are reused when they have the right shape
        v, i = self.currentMesh.vertices, self.currentMesh.indices
        if not (isinstance(v, np.ndarray) and v.dtype == np.float32 and v.size == (R + 1) * (R + 1) * 12):
            self.currentMesh.vertices = np.empty(((R + 1) * (R + 1), 12), np.float32)
        if not (isinstance(i, np.ndarray) and i.dtype == np.uint32 and i.size == 6 * R * R):
            self.currentMesh.indices = np.empty(6 * R * R, np.uint32)
        self.currentMesh.bounds = ((0.5 * d.tileSize, 0.5 * d.tileHeight, 0.5 * d.tileSize),
                                   (d.tileSize, d.tileHeight, d.tileSize))
        self._native(requirements, lambda: _h.heightmap_mesh(self.meshType, self.currentMesh.vertices, self.currentMesh.indices, R,
                                                             d.inputResolution, d.marginPix, d.tileHeight, d.tileSize, d.data))


# ---- pipeline executor (the scheduling chain of BasePipeline) ------------------------------------------
class BasePipeline:
    """Queue + stage chaining of Pipeline/Executable/Pipeline.cs: stage i's OnStageScheduledAction feeds
    stage i+1's ReceiveHandledInput with the JobHandle as dependency (:130-151); Update() schedules the
    next queued item (:154-158,216-230), LateUpdate() completes it and fires completeAction (:160-181)."""

    def __init__(self, stages, alias="pipeline", contextManager=None):
        if not stages:
            raise Exception("No stages in pipeline")
        self.alias = alias
        self.contextManager = contextManager      # PipelineStateManager handed to every work item (Pipeline.cs:95-103)
        self.dependencyHell = []
        self.stage_instances = list(stages)
        self.queue = deque()
        self.pipelineRunning = False
        self.pipelineQueued = False
        self.pipelineHandle = None
        self.activeItem = None
        for i, st in enumerate(self.stage_instances):
            if i + 1 < len(self.stage_instances):
                st.OnStageScheduledAction = self.stage_instances[i + 1].ReceiveHandledInput
            else:
                st.OnStageScheduledAction = self.OnPipelineFullyScheduled

    def Enqueue(self, input, scheduleAction=None, completeAction=None, dependency=None):
        self.queue.append(PipelineWorkItem(input, completeAction, scheduleAction, dependency, self.contextManager))

    def WorkIsSchedulable(self, item):         # Pipeline.cs:256-265
        return all([st.IsSchedulable(item) for st in self.stage_instances])

    def GetNextJob(self):
        """Pipeline.cs:183-215: items whose stages are not schedulable yet (a context buffer that does not exist, a locked
        one) wait in `dependencyHell` and are retried before the queue on every Update."""
        for i, item in enumerate(list(self.dependencyHell)):
            if self.WorkIsSchedulable(item):
                self.dependencyHell.remove(item)
                return item
        while self.queue:
            item = self.queue.popleft()
            if self.WorkIsSchedulable(item):
                return item
            self.dependencyHell.append(item)
        return None

    def Schedule(self, workItem):
        if self.pipelineRunning:
            raise Exception("Pipeline already running")
        self.activeItem = workItem
        self.pipelineQueued = True
        self.stage_instances[0].ReceiveHandledInput(workItem, workItem.dependency)

    def OnPipelineFullyScheduled(self, requirements, dependency):
        self.pipelineHandle = dependency
        self.pipelineQueued = False
        self.pipelineRunning = True
        if requirements.scheduledAction is not None:
            requirements.scheduledAction(requirements.data, dependency)

    def Update(self):
        if not self.pipelineRunning and not self.pipelineQueued:
            job = self.GetNextJob()
            if job is not None:
                self.Schedule(job)

    def LateUpdate(self):
        if self.pipelineRunning:
            self.pipelineHandle.Complete()
            self.pipelineRunning = False
            for st in self.stage_instances:
                st.OnStageComplete()
            item, self.activeItem = self.activeItem, None
            if item.completeAction is not None:
                item.completeAction(item.data)

    def Run(self, input, **kw):
        """Convenience: Enqueue + pump Update/LateUpdate once (what the editor window does per frame)."""
        self.Enqueue(input, **kw)
        self.Update()
        self.LateUpdate()
        return input

    def Destroy(self):
        for st in self.stage_instances:
            st.Destroy()


# ---- PipelineState: named buffers that outlive a pipeline run, on the GPU -------------------------------------------------
class PipelineStateManager:
    """Mirror of the float-buffer part of PipelineStateManager (Pipeline/PipelineState/PipelineStateManager.cs:39-160) with the
    buffers living in HBM (nz_context_*): GetBuffer / BufferExists / ReleaseBuffer / IsLocked / TrySetLock /
    SaveBufferToDisk, and the saved-state load of GetBuffer (:64-72) through the reference's own on-disk format (serde.py).
    A buffer is referred to by its NAME; there is no host array behind it until someone downloads it."""

    def __init__(self):
        self.savedState = None
        self._locks = {}

    def SetSavePath(self, path, saveName, saveVersion):
        from . import serde
        self.savedState = serde.PipelineSerdeManager(path, saveName, saveVersion)

    def GetBuffer(self, name, size=-1, ignoreSaved=False):
        """Returns the buffer's name once it exists on the device; a saved copy on disk is uploaded the first time."""
        if _h.context_exists(name) < 0 and self.savedState is not None and not ignoreSaved:
            n = self.savedState.CachedSize(name)
            if n > 0:
                _h.context_upload(name, self.savedState.ReadData(name, np.float32, n))
        return name

    def BufferExists(self, name):
        return _h.context_exists(name) >= 0

    def ReleaseBuffer(self, name):
        existed = self.BufferExists(name)
        _h.context_release(name)
        self._locks.pop(name, None)
        return existed

    def IsLocked(self, key):
        h = self._locks.get(key)
        return h is not None and not h.IsCompleted

    def TrySetLock(self, key, handle, spyHandle=None):
        if self.IsLocked(key):
            return False
        self._locks[key] = handle
        return True

    def SaveBufferToDisk(self, name, size=-1):
        if self.savedState is None:
            raise ValueError("No serde manager is active")
        self.savedState.WriteData(_h.context_download(name), name)


def _context_buffer_name(d, alias):
    return f"{d.xpos}_{d.zpos}__{d.resolution}__{alias}"          # WriteGeneratorContextStage.getBufferName, :21-23


class WriteGeneratorContextStage(GpuStage):
    """PipelineState/Stage/WriteGeneratorContextStage.cs:13-55: parks the work item's tile in the state manager under
    "{xpos}_{zpos}__{resolution}__{contextAlias}".  On the GPU that is one device-to-device copy out of the resident tile."""

    def __init__(self, contextAlias=""):
        super().__init__()
        self.contextAlias = contextAlias

    def IsSchedulable(self, job):
        if job.stageManager is None:
            return False
        return not job.stageManager.IsLocked(_context_buffer_name(job.data, self.contextAlias))

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        gd = requirements.data
        name = requirements.stageManager.GetBuffer(_context_buffer_name(gd, self.contextAlias), gd.resolution * gd.resolution,
                                                   ignoreSaved=True)
        self._native(requirements, lambda: _h.context_write(name, gd.data))
        requirements.stageManager.TrySetLock(name, self.jobHandle, self.jobHandle)


class ReadGeneratorContextStage(GpuStage):
    """PipelineState/Stage/ReadGeneratorContextStage.cs:13-53: the work item's tile := the named buffer."""

    def __init__(self, contextAlias=""):
        super().__init__()
        self.contextAlias = contextAlias

    def IsSchedulable(self, job):
        if job.stageManager is None:
            return False
        name = _context_buffer_name(job.data, self.contextAlias)
        job.stageManager.GetBuffer(name)          # a saved copy on disk counts as existing
        if not job.stageManager.BufferExists(name):
            return False
        return not job.stageManager.IsLocked(name)

    def Schedule(self, requirements, dependency):
        self.CheckRequirements(GeneratorData, requirements)
        gd = requirements.data
        name = requirements.stageManager.GetBuffer(_context_buffer_name(gd, self.contextAlias), gd.resolution * gd.resolution)
        self._native(requirements, lambda: _h.context_read(name, gd.data))


# ---- ReducePipeline: two upstream pipelines joined by a reduce pipeline ------------------------------------------------------
class ReducePipeline(BasePipeline):
    """Pipeline/Executable/ReducePipeline.cs:32-165: takes ONE work item, requests the same item (same uuid, position and
    resolution; the right side on a buffer of its own) from both upstream pipelines, and when both have completed runs its
    own stages on a ReduceData{data = left, rightData = right}.

    On the GPU the join needs no host traffic: the reference gives both upstream items the SAME uuid (:104-112), and
    GpuResidency keys the residency scope by uuid, so with `keepResident` on the last stage of each upstream pipeline both
    operands are still in HBM when this pipeline's ReduceStage runs, and the scope closes (one download of the result) at
    the end of THIS pipeline.  The C# ReducePipeline works unchanged with Gpu* stages for the same reason."""

    def __init__(self, stages, upstreamPipelineLeft, upstreamPipelineRight, alias="reduce", contextManager=None):
        super().__init__(stages, alias, contextManager)
        self.upstreamPipelineLeft, self.upstreamPipelineRight = upstreamPipelineLeft, upstreamPipelineRight
        self.upstreamsRunning = False
        self.currentWorkItem = None
        self.rightData = None
        self.currentDataLength = 0

    def GetDependencies(self):
        deps = [self.upstreamPipelineLeft, self.upstreamPipelineRight, self]
        for p in (self.upstreamPipelineLeft, self.upstreamPipelineRight):
            if hasattr(p, "GetDependencies"):
                deps += p.GetDependencies()
        return deps

    def Update(self):          # OnUpdate, :64-80
        if not self.pipelineRunning and not self.pipelineQueued and not self.upstreamsRunning and self.queue:
            self.upstreamsRunning = True
            self.ScheduleUpstreams(self.queue.popleft())
        for p in (self.upstreamPipelineLeft, self.upstreamPipelineRight):
            p.Update()

    def LateUpdate(self):
        for p in (self.upstreamPipelineLeft, self.upstreamPipelineRight):
            p.LateUpdate()
        super().LateUpdate()

    def ScheduleUpstreams(self, wi):     # :82-122
        left = wi.data
        if left.data.size != self.currentDataLength or self.rightData is None:
            self.currentDataLength = left.data.size
            self.rightData = np.empty(self.currentDataLength, np.float32)
        right = GeneratorData(left.uuid, self.rightData, left.resolution, left.xpos, left.zpos)
        self.currentWorkItem = {"status": {"L": False, "R": False}, "stages": {"L": left, "R": right}, "action": wi.completeAction}
        self.upstreamPipelineLeft.Enqueue(left, completeAction=lambda res: self.OnCompleteUpstream(res, "L"))
        self.upstreamPipelineRight.Enqueue(right, completeAction=lambda res: self.OnCompleteUpstream(res, "R"))

    def OnCompleteUpstream(self, res, side):     # :124-150
        w = self.currentWorkItem
        w["status"][side] = True
        w["stages"][side] = res
        if all(w["status"].values()):
            self.upstreamsRunning = False
            d = res
            rd = ReduceData(d.uuid, w["stages"]["L"].data, w["stages"]["R"].data, d.resolution, d.xpos, d.zpos)
            self.Schedule(PipelineWorkItem(rd, w["action"], None, None, self.contextManager))

    def Run(self, input, **kw):
        """Enqueue and pump frames until the joined item has completed."""
        done = []
        user = kw.pop("completeAction", None)

        def finished(d):
            done.append(d)
            if user is not None:
                user(d)
        self.Enqueue(input, completeAction=finished, **kw)
        for _ in range(8):
            self.Update()
            self.LateUpdate()
            if done:
                break
        if not done:
            raise Exception(f"{self.alias}: the joined work item did not complete")
        return done[0]
