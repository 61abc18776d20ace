"""noize-job_b200 — B200-native (sm_100a) heightmap hot path of xshazwar/noize-job.

    lib      ctypes loader of libnoize_b200.so (the C ABI of include/noize_b200.h)
    host     host layer  (nz_*):     host buffers, one call per reference job delegate
    device   device layer (nz_dev_*): device buffers on a stream, rectangular row bands
    stages   mirror of the reference's PipelineStage classes bound to the host layer
    bands    row-band partition of one large heightmap across the GPUs of a box
    tiles    tile sharding of a tiled world across the GPUs of a box (independent tiles, several in flight per GPU)

The directory name carries a hyphen; import it as `noize_job_b200` (see /noize_job_b200.py).
"""
from . import lib, host, device, stages  # noqa: F401
from .lib import NzError, load, build  # noqa: F401
from .stages import (FractalNoise, KernelFilterType, GaussSigma, MeshType, JobHandle, StageIO, GeneratorData,  # noqa: F401
                     MeshStageData, Mesh, PipelineWorkItem, PipelineStage, NoiseStage, KernelFilterStage,
                     StageGaussianBlur, StageSmoothBlur, ErosionFilterStage, FlowMapStage, MeshTileStage, BasePipeline,
                     ConstantOperationType, ReductionType, ReduceData, DownsampleData, StageThermalErosion, ErosionStageSubtractiveFlow,
                     ConstantStage, ReduceStage, CurveStage, CropStage, GpuStage, GpuResidency,
                     PipelineStateManager, WriteGeneratorContextStage, ReadGeneratorContextStage, ReducePipeline)
