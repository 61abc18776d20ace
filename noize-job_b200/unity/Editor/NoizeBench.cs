// NoizeBench.cs — Editor-side harness for the day a Unity box exists.  NOT compiled or run in this repository's image (no
// Unity / mono / dotnet there); until it has run, parity of the oracle against the real Burst path stays UNPINNED.
//
//   Unity -batchmode -nographics -projectPath <project with xshazwar.noize + this folder> \
//         -executeMethod xshazwar.noize.interop.b200.editor.NoizeBench.Run -quit -logFile bench.log \
//         [-noizeOut <dir>] [-noizeConfigs C1,C2,C3] [-noizeReps 5]
//
// For each BASELINE config it runs the chain twice — on the reference's Burst stages and on the Gpu* stages of
// Interop/NoizeB200.cs — with identical parameters, and for both
//   * dumps the tile after EVERY stage through the reference's own PipelineSerdeManager
//     (Pipeline/PipelineState/PipelineSerialization.cs:210-228): <out>/save__burst_C2/data/<stage>.data and
//     <out>/save__gpu_C2/data/<stage>.data, raw little-endian f32 — the format tools/dump_chain.py and serde.py write, so
//     tools/compare_dump.py diffs any two of {Burst, GPU-in-Unity, GPU-from-Python, oracle} byte for byte;
//   * logs wall-clock milliseconds per stage and per chain the way the pipeline does (Stopwatch around schedule ->
//     JobHandle.Complete, Pipeline/Executable/Pipeline.cs:116,126,170-171), best and mean of `reps`, together with
//     JobsUtility.JobWorkerCount and SystemInfo.processorCount: the CPU baseline BASELINE.md section 3 asks for.
// Stages are driven exactly as BasePipeline drives them: stage.ReceiveHandledInput(workItem, dependency) with the previous
// stage's handle (Pipeline.cs:130-151), Complete(), then OnStageComplete() (Pipeline.cs:232-240).
//
// Value erosion has no PipelineStage in the reference tree (ErosionKernelJob is declared but unbound, KernelJob.cs:317-350),
// so the Burst arm calls the job's static Schedule directly, `iterations` times, as a stage would.
#if UNITY_EDITOR
using System;
using System.Collections.Generic;
using System.Diagnostics;
using System.Globalization;
using System.IO;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Jobs;
using Unity.Jobs.LowLevel.Unsafe;
using UnityEngine;

using xshazwar.noize.pipeline;
using xshazwar.noize.filter;
using xshazwar.noize.generate;
using xshazwar.noize.geologic;
using xshazwar.noize.mesh;

namespace xshazwar.noize.interop.b200.editor {

    public static unsafe class NoizeBench {

        class Step {
            public string name;
            public Func<PipelineWorkItem, JobHandle, JobHandle> schedule;   // returns the handle to complete
            public Action complete;                                          // OnStageComplete
        }

        static Step FromStage(string name, PipelineStage stage) {
            JobHandle last = default;
            stage.OnStageScheduledAction = (wi, h) => { last = h; };
            return new Step {
                name = name,
                schedule = (wi, dep) => { stage.ReceiveHandledInput(wi, dep); return last; },
                complete = stage.OnStageComplete
            };
        }

        static T Make<T>(Action<T> init) where T : ScriptableObject {
            T s = ScriptableObject.CreateInstance<T>();
            init(s);
            return s;
        }

        // ---- the chains of BASELINE.json configs[0..2] -------------------------------------------------------------
        static List<Step> BurstChain(string cfg) {
            var noiseType = cfg == "C3" ? NoiseStage.FractalNoise.Cellular : NoiseStage.FractalNoise.Simplex;
            var steps = new List<Step> {
                FromStage("noise", Make<NoiseStage>(s => { s.noiseType = noiseType; s.hurst = 0.4f; s.startingAmplitude = 1f; s.octaves = 13;
                                                           s.stepdown = 2f; s.detuneRate = 0f; s.noiseSize = 1700; }))
            };
            if (cfg == "C1") return steps;
            steps.Add(FromStage("gauss5x17", Make<KernelFilterStage>(s => { s.filter = KernelFilterType.Gauss5_S1; s.iterations = 17; })));
            steps.Add(FromStage("flow5", Make<FlowMapStage>(s => { s.iterations = 5; s.normMin = 0f; s.normMax = 0.005f; })));
            if (cfg == "C2")
                steps.Add(new Step {
                    name = "erosion5",
                    schedule = (wi, dep) => {
                        GeneratorData d = (GeneratorData) wi.data;
                        JobHandle h = dep;
                        for (int i = 0; i < 5; i++) h = ErosionKernelJob.Schedule(d.data, d.resolution, h);
                        return h;
                    },
                    complete = () => { }
                });
            return steps;
        }

        static List<Step> GpuChain(string cfg) {
            var noiseType = cfg == "C3" ? NoiseStage.FractalNoise.Cellular : NoiseStage.FractalNoise.Simplex;
            var steps = new List<Step> {
                FromStage("noise", Make<GpuNoiseStage>(s => { s.noiseType = noiseType; s.hurst = 0.4f; s.startingAmplitude = 1f; s.octaves = 13;
                                                              s.stepdown = 2f; s.detuneRate = 0f; s.noiseSize = 1700; }))
            };
            if (cfg == "C1") return steps;
            steps.Add(FromStage("gauss5x17", Make<GpuKernelFilterStage>(s => { s.filter = KernelFilterType.Gauss5_S1; s.iterations = 17; })));
            steps.Add(FromStage("flow5", Make<GpuFlowMapStage>(s => { s.iterations = 5; s.normMin = 0f; s.normMax = 0.005f; })));
            if (cfg == "C2") steps.Add(FromStage("erosion5", Make<GpuErosionFilterStage>(s => { s.iterations = 5; })));
            return steps;
        }

        static int Resolution(string cfg) => cfg == "C1" ? 256 : cfg == "C2" ? 1024 : 4096;

        // Runs the chain stage by stage (each stage completed before the next is scheduled, so a per-stage dump and time
        // exist), then once more as ONE dependency chain completed at the end (what the pipeline does) for the chain time.
        static void RunArm(string arm, string cfg, List<Step> steps, string outDir, int reps) {
            int res = Resolution(cfg);
            var tile = new NativeArray<float>(res * res, Allocator.Persistent, NativeArrayOptions.ClearMemory);
            var serde = new PipelineSerdeManager(outDir, $"{arm}_{cfg}", "noize-bench-1");
            var perStage = new Dictionary<string, List<double>>();
            var chain = new List<double>();
            try {
                for (int rep = 0; rep < reps + 1; rep++) {            // rep 0 is the warm-up (Burst compiles, CUDA context)
                    var wi = new PipelineWorkItem { data = new GeneratorData { uuid = $"{arm}-{cfg}", data = new NativeSlice<float>(tile), resolution = res, xpos = 0, zpos = 0 } };
                    foreach (Step st in steps) {
                        var sw = Stopwatch.StartNew();
                        JobHandle h = st.schedule(wi, default);
                        h.Complete();
                        st.complete();
                        sw.Stop();
                        if (rep > 0) {
                            if (!perStage.ContainsKey(st.name)) perStage[st.name] = new List<double>();
                            perStage[st.name].Add(sw.Elapsed.TotalMilliseconds);
                        }
                        if (rep == reps) serde.WriteData<float>(tile.GetUnsafeReadOnlyPtr(), tile.Length * sizeof(float), st.name, tile.Length);
                    }
                    // the same chain as one dependency graph
                    wi = new PipelineWorkItem { data = new GeneratorData { uuid = $"{arm}-{cfg}-chain", data = new NativeSlice<float>(tile), resolution = res, xpos = 0, zpos = 0 } };
                    var wall = Stopwatch.StartNew();
                    JobHandle dep = default;
                    foreach (Step st in steps) dep = st.schedule(wi, dep);
                    dep.Complete();
                    foreach (Step st in steps) st.complete();
                    wall.Stop();
                    if (rep > 0) chain.Add(wall.Elapsed.TotalMilliseconds);
                    if (rep == reps) serde.WriteData<float>(tile.GetUnsafeReadOnlyPtr(), tile.Length * sizeof(float), "chain", tile.Length);
                }
            } finally {
                tile.Dispose();
            }
            double cells = (double) res * res;
            foreach (var kv in perStage) Log(arm, cfg, kv.Key, kv.Value, cells);
            Log(arm, cfg, "chain", chain, cells);
        }

        static void Log(string arm, string cfg, string what, List<double> ms, double cells) {
            double best = double.MaxValue, sum = 0;
            foreach (double v in ms) { best = Math.Min(best, v); sum += v; }
            double mean = sum / ms.Count;
            // one machine-readable line per measurement (tools/compare_dump.py --log collects them)
            UnityEngine.Debug.Log(string.Format(CultureInfo.InvariantCulture,
                "NOIZE_BENCH {{\"arm\":\"{0}\",\"config\":\"{1}\",\"stage\":\"{2}\",\"ms_best\":{3:F4},\"ms_mean\":{4:F4},\"mcells_s_best\":{5:F2},\"reps\":{6},\"job_workers\":{7},\"processors\":{8}}}",
                arm, cfg, what, best, mean, cells / best / 1e3, ms.Count, JobsUtility.JobWorkerCount, SystemInfo.processorCount));
        }

        static string Arg(string name, string fallback) {
            string[] a = Environment.GetCommandLineArgs();
            for (int i = 0; i + 1 < a.Length; i++) if (a[i] == name) return a[i + 1];
            return fallback;
        }

        public static void Run() {
            string outDir = Arg("-noizeOut", Path.Combine(Application.dataPath, "..", "noize_bench_out"));
            string[] cfgs = Arg("-noizeConfigs", "C1,C2,C3").Split(',');
            int reps = int.Parse(Arg("-noizeReps", "5"), CultureInfo.InvariantCulture);
            UnityEngine.Debug.Log($"NOIZE_BENCH start: out={outDir} job workers={JobsUtility.JobWorkerCount} processors={SystemInfo.processorCount} native={System.Runtime.InteropServices.Marshal.PtrToStringAnsi(Native.nz_version())}");
            foreach (string cfg in cfgs) {
                RunArm("burst", cfg, BurstChain(cfg), outDir, reps);
                RunArm("gpu", cfg, GpuChain(cfg), outDir, reps);
            }
            GpuResidency.CloseAll();
            UnityEngine.Debug.Log("NOIZE_BENCH done");
        }
    }
}
#endif
