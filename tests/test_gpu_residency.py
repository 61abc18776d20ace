"""Residency owned by the STAGE OBJECTS (the C# GpuStage logic, mirrored in stages.py and host_cpp/noize_stages.hpp):
the first GPU stage of a work item opens a scope keyed by StageIO.uuid, the last one before a host consumer closes it.
No test here opens a scope around the pipeline: that is the point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 256


def _gen_stages(nz):
    return [nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700),
            nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=4),
            nz.FlowMapStage(iterations=3, normMin=0.0, normMax=0.005),
            nz.ErosionFilterStage(iterations=2)]


def _reference(nz):
    a = np.zeros(N * N, np.float32)
    nz.host.fractal(a, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)     # stand-alone calls: H2D + D2H around every stage
    nz.host.kernel_filter(a, None, 2, N, 4)
    nz.host.flowmap(a, N, 3, 0.0, 0.005)
    nz.host.min_erosion(a, N, 2)
    return a


class HostStage:
    """A stage that reads HOST memory when it is scheduled, like a Burst stage would when its job runs."""

    def __init__(self, nz):
        base = nz.PipelineStage
        outer = self

        class _Stage(base):
            def Schedule(self, requirements, dependency):
                outer.seen = requirements.data.data.copy()
                requirements.data.data *= np.float32(0.5)           # ... and writes it
                self.jobHandle = nz.JobHandle()
        self.stage = _Stage()
        self.seen = None


def test_stages_own_the_scope_and_one_download_ends_the_chain(nz):
    data = np.zeros(N * N, np.float32)
    scheduled = []
    pipe = nz.BasePipeline(_gen_stages(nz))
    pipe.Enqueue(nz.GeneratorData("tile-a", data, N, 0, 0), scheduleAction=lambda d, h: scheduled.append(nz.GpuResidency.IsOpen("tile-a")))
    pipe.Update()
    # the fully-scheduled hook is a host consumer: the last GPU stage closed the scope before handing its handle on
    assert scheduled == [False]
    pipe.LateUpdate()
    assert not nz.host.thread_in_scope()
    assert np.array_equal(data, _reference(nz))


def test_scope_is_shared_by_consecutive_gpu_stages(nz):
    """Between two GPU stages the tile stays in HBM: the host slice is NOT written until the chain ends."""
    data = np.full(N * N, -1.0, np.float32)
    stages = _gen_stages(nz)
    seen_between = {}

    class Spy(nz.KernelFilterStage):             # a GPU stage that looks at host memory before it runs
        def Schedule(self, requirements, dependency):
            seen_between["open"] = nz.GpuResidency.IsOpen(requirements.data.uuid)
            seen_between["host_untouched"] = bool((requirements.data.data == -1.0).all())
            super().Schedule(requirements, dependency)
    stages[1] = Spy(nz.KernelFilterType.Gauss5_S1, iterations=4)
    nz.BasePipeline(stages).Run(nz.GeneratorData("tile-b", data, N, 0, 0))
    assert seen_between == {"open": True, "host_untouched": True}
    assert np.array_equal(data, _reference(nz))


def test_host_stage_in_the_middle_sees_current_data(nz):
    """GPU, GPU, HOST, GPU: the scope closes in front of the host stage (it reads the filtered noise) and a new one opens
    after it (the flow map uploads what the host stage wrote)."""
    data = np.zeros(N * N, np.float32)
    st = _gen_stages(nz)
    probe = HostStage(nz)
    nz.BasePipeline([st[0], st[1], probe.stage, st[2], st[3]]).Run(nz.GeneratorData("tile-c", data, N, 0, 0))
    want = np.zeros(N * N, np.float32)
    nz.host.fractal(want, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
    nz.host.kernel_filter(want, None, 2, N, 4)
    assert np.array_equal(probe.seen, want)
    want *= np.float32(0.5)
    nz.host.flowmap(want, N, 3, 0.0, 0.005)
    nz.host.min_erosion(want, N, 2)
    assert np.array_equal(data, want)
    assert not nz.GpuResidency.IsOpen("tile-c")


def test_keep_resident_hands_the_tile_to_the_next_pipeline(nz, oracle):
    """Generator pipeline -> mesh pipeline on the same uuid (Scripts/MeshTileGenerator.cs:94-138): with keepResident the
    heightmap is downloaded once (the handle's contract) and the mesh stage finds it in HBM."""
    R = N - 8
    data = np.zeros(N * N, np.float32)
    st = _gen_stages(nz)
    st[-1].keepResident = True
    nz.BasePipeline(st).Run(nz.GeneratorData("tile-d", data, N, 0, 0))
    assert nz.GpuResidency.IsOpen("tile-d")
    want = _reference(nz)
    assert np.array_equal(data, want)
    data_before = data.copy()
    data[:] = 7.0                                # prove residency: the mesh must come from the resident copy, not from this
    mesh = nz.MeshStageData("tile-d", data, R, N, 4, R * (500.0 / 256.0), 2000.0)
    nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)]).Run(mesh)
    assert not nz.GpuResidency.IsOpen("tile-d")
    rv, ri = oracle.heightmap_mesh(1, data_before.reshape(N, N), R, 4, 2000.0, R * (500.0 / 256.0))
    assert np.array_equal(mesh.mesh.indices, ri)
    assert np.abs(mesh.mesh.vertices - rv).max() <= 1e-5 * 2000.0


def test_failed_stage_closes_the_scope(nz):
    data = np.zeros(N * N, np.float32)
    st = _gen_stages(nz)
    st[1].filter = 99                            # rejected by the native call with NZ_E_INVALID
    with pytest.raises(nz.NzError):
        nz.BasePipeline(st).Run(nz.GeneratorData("tile-e", data, N, 0, 0))
    assert not nz.GpuResidency.IsOpen("tile-e")
    assert not nz.host.thread_in_scope()


def test_blur_stages_match_oracle(nz, oracle):
    """GpuStageGaussianBlur / GpuStageSmoothBlur (StageGaussianBlur.cs:33-46, StageSmoothBlur.cs): sigma table x width."""
    rng = np.random.default_rng(20221018)
    src = rng.random((N, N), dtype=np.float32)
    for sigma, width, it in ((nz.GaussSigma.s1d50, 7, 2), (nz.GaussSigma.s4d00, 25, 1), (nz.GaussSigma.s0d50, 4, 3)):
        d = src.copy().reshape(-1)
        nz.BasePipeline([nz.StageGaussianBlur(iterations=it, sigma=sigma, width=width)]).Run(nz.GeneratorData("blur", d, N))
        assert np.abs(d.reshape(N, N) - oracle.gauss_filter(src, width, sigma, it)).max() <= 1e-6
    d = src.copy().reshape(-1)
    nz.BasePipeline([nz.StageSmoothBlur(iterations=2, width=5)]).Run(nz.GeneratorData("blur", d, N))
    assert np.abs(d.reshape(N, N) - oracle.smooth_filter(src, 5, 2)).max() <= 1e-6


def test_context_stages_keep_tiles_on_the_device_between_pipelines(nz, tmp_path):
    """Write/ReadGeneratorContextStage + PipelineStateManager (PipelineState/Stage/*.cs, PipelineStateManager.cs:39-160) with
    the named buffers in HBM: a generator pipeline parks its tile, a second pipeline picks it up; the buffer round-trips
    through the reference's on-disk format."""
    mgr = nz.PipelineStateManager()
    gen = nz.BasePipeline(_gen_stages(nz)[:2] + [nz.WriteGeneratorContextStage("terrain")], contextManager=mgr)
    use = nz.BasePipeline([nz.ReadGeneratorContextStage("terrain"), nz.FlowMapStage(iterations=3, normMin=0.0, normMax=0.005)],
                          contextManager=mgr)
    name = f"0_424__{N}__terrain"
    out = np.zeros(N * N, np.float32)
    done = []
    # the consumer is enqueued FIRST: its buffer does not exist yet, so the item waits (Pipeline.cs:200-214)
    use.Enqueue(nz.GeneratorData("consumer", out, N, 0, 424), completeAction=lambda d: done.append(d.uuid))
    use.Update()
    assert not use.pipelineRunning and len(use.dependencyHell) == 1 and not mgr.BufferExists(name)
    tile = np.zeros(N * N, np.float32)
    gen.Run(nz.GeneratorData("producer", tile, N, 0, 424))
    assert mgr.BufferExists(name) and nz.host.context_exists(name) == N * N
    want = np.zeros(N * N, np.float32)
    nz.host.fractal(want, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 424, 1700)
    nz.host.kernel_filter(want, None, 2, N, 4)
    assert np.array_equal(tile, want) and np.array_equal(nz.host.context_download(name), want)
    use.Update()
    use.LateUpdate()
    assert done == ["consumer"] and not use.dependencyHell
    nz.host.flowmap(want, N, 3, 0.0, 0.005)
    assert np.array_equal(out, want)
    # SaveBufferToDisk -> a fresh manager finds the saved copy and uploads it on first use (PipelineStateManager.cs:64-72)
    mgr.SetSavePath(str(tmp_path), "world", "1")
    mgr.SaveBufferToDisk(name)
    assert mgr.ReleaseBuffer(name) and not mgr.BufferExists(name)
    mgr2 = nz.PipelineStateManager()
    mgr2.SetSavePath(str(tmp_path), "world", "1")
    out2 = np.zeros(N * N, np.float32)
    nz.BasePipeline([nz.ReadGeneratorContextStage("terrain")], contextManager=mgr2).Run(nz.GeneratorData("reload", out2, N, 0, 424))
    tile_only = np.zeros(N * N, np.float32)
    nz.host.fractal(tile_only, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 424, 1700)
    nz.host.kernel_filter(tile_only, None, 2, N, 4)
    assert np.array_equal(out2, tile_only)
    mgr2.ReleaseBuffer(name)
    with pytest.raises(nz.NzError) as e:
        nz.host.context_read("no-such-buffer", out2)
    assert e.value.code == nz.lib.NZ_E_STATE


def test_reduce_pipeline_joins_two_upstreams_on_the_device(nz):
    """ReducePipeline (Pipeline/Executable/ReducePipeline.cs:32-165): both upstream items carry the same uuid, so with
    keepResident on the upstreams' last stages the two operands meet in HBM — the host copies are poisoned before the reduce
    runs, so a result that matches can only have come from the resident mirrors."""
    left_stages = [nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700),
                   nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=4)]
    right_stages = [nz.NoiseStage(nz.FractalNoise.Cellular, hurst=0.7, octaves=5, noiseSize=300),
                    nz.ConstantStage(nz.ConstantOperationType.MULTIPLY, 0.5)]
    left_stages[-1].keepResident = True
    right_stages[-1].keepResident = True
    left, right = nz.BasePipeline(left_stages, "left"), nz.BasePipeline(right_stages, "right")
    poisoned = []

    class Poison(nz.ReduceStage):          # first stage of the joined pipeline: both host operands are current, and get poisoned
        def Schedule(self, requirements, dependency):
            d = requirements.data
            poisoned.append((d.data.copy(), d.rightData.copy(), nz.GpuResidency.IsOpen(d.uuid)))
            d.data[:] = -5.0
            d.rightData[:] = -7.0
            super().Schedule(requirements, dependency)
    rp = nz.ReducePipeline([Poison(nz.ReductionType.MAX), nz.ConstantStage(nz.ConstantOperationType.MULTIPLY, 2.0)], left, right)
    data = np.zeros(N * N, np.float32)
    out = rp.Run(nz.GeneratorData("joined", data, N, 0, 424))
    a = np.zeros(N * N, np.float32)
    nz.host.fractal(a, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 424, 1700)
    nz.host.kernel_filter(a, None, 2, N, 4)
    b = np.zeros(N * N, np.float32)
    nz.host.fractal(b, N, 5, 0.7, 1.0, 2.0, 0.0, 5, 0, 424, 300)
    nz.host.constant(b, None, 0, 0.5, N)
    la, ra, was_open = poisoned[0]
    assert was_open and np.array_equal(la, a) and np.array_equal(ra, b)          # keepResident flushed both, the scope stayed
    assert np.array_equal(data, np.maximum(a, b) * np.float32(2.0))
    assert isinstance(out, nz.GeneratorData)                                       # ReduceStage.TransformData (ReduceStage.cs:52-61)
    assert not nz.GpuResidency.IsOpen("joined")


def test_scope_closed_under_a_thread_still_inside_releases_late_mirrors(nz):
    """A scope closed by one thread while another is still inside it: the second thread's later stage calls keep working on
    the scope object it entered, and what they leave behind is released when that thread leaves (no device memory leak, and the
    library keeps working)."""
    import threading
    import numpy as np
    res = 256
    sid = nz.host.scope_create()
    entered, closed, done = threading.Event(), threading.Event(), threading.Event()
    err = []

    def worker():
        try:
            nz.host.scope_enter(sid)
            entered.set()
            closed.wait(10)
            a = np.zeros(res * res, np.float32)
            nz.host.fractal(a, res, 3, 0.4, 1.0, 2.0, 0.0, 3, 0, 0, 300)     # lands in the closed scope's object
            nz.host.scope_leave()                                            # last reference: the mirror is released here
        except Exception as e:                                               # noqa: BLE001
            err.append(e)
        finally:
            done.set()

    t = threading.Thread(target=worker)
    t.start()
    assert entered.wait(10)
    nz.host.scope_close(sid)
    closed.set()
    assert done.wait(30)
    t.join()
    assert not err, err
    b = np.zeros(res * res, np.float32)
    nz.host.fractal(b, res, 3, 0.4, 1.0, 2.0, 0.0, 3, 0, 0, 300)
    assert float(b.max()) > 0.0
