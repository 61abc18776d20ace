"""Extract the reference's literal constants into tests/golden/ (run in the build container only).

The reference ships no tests or golden vectors; the only numeric ground truth it holds for the hot
path are its literal tables.  This script parses them out of /root/reference (read-only, absent
on the GPU box) and writes small fixtures that travel with the repo:

  kernel_tables.npz
     blur/<sigma_index>/<width>  : every Gaussian table of Filter/Kernel/Blur/BlurKernels.cs:59-316
     kj/<name>                   : gauss*_s*, smooth3, sobel3_*, prewitt3_* of Filter/Kernel/KernelJob.cs:97-136
  stage_params.json              : parameter values of BasicDemo~/*.asset + DynamicNoise.unity used by the configs

    python tests/golden/make_golden.py
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def floats(body):
    return np.array([float(v.strip().rstrip("f")) for v in body.replace("\n", " ").split(",") if v.strip()],
                    dtype=np.float64).astype(np.float32)


def main():
    out = {}
    src = open(os.path.join(REF, "Filter/Kernel/Blur/BlurKernels.cs")).read()
    names = re.findall(r"^\s+(s\dd\d\d),?\s*$", src.split("public static class BlurHelper")[0], flags=re.M)
    pos = [(m.start(), m.group(1)) for m in re.finditer(r"GaussSigma\.(s\dd\d\d)\s*, new List", src)]
    pos.append((len(src), None))
    for (a, name), (b, _) in zip(pos[:-1], pos[1:]):
        for body in re.findall(r"new float\[\] \{([^}]*)\}", src[a:b]):
            t = floats(body)
            out[f"blur/{names.index(name)}/{t.size}"] = t
    kj = open(os.path.join(REF, "Filter/Kernel/KernelJob.cs")).read()
    for m in re.finditer(r"public static float\[\] (\w+)\s*=\s*\{([^}]*)\}", kj):
        out[f"kj/{m.group(1)}"] = floats(m.group(2))
    for m in re.finditer(r"public static float (\w+Factor)\s*=\s*([^;]*);", kj):
        expr = m.group(2).replace("f", "").strip()
        out[f"kj/{m.group(1)}"] = np.array([np.float32(eval(expr))], np.float32)
    np.savez(os.path.join(HERE, "kernel_tables.npz"), **out)

    def asset(name, keys):
        txt = open(os.path.join(REF, "BasicDemo~", name)).read()
        return {k: float(re.search(rf"^\s+{k}: (\S+)", txt, flags=re.M).group(1)) for k in keys}

    params = {
        "Simplex.asset": asset("Simplex.asset", ["noiseType", "hurst", "startingAmplitude", "octaves", "stepdown", "detuneRate", "noiseSize"]),
        "GaussHF.asset": asset("GaussHF.asset", ["filter", "iterations"]),
        "GaussLF.asset": asset("GaussLF.asset", ["filter", "iterations"]),
        "Sobel2D.asset": asset("Sobel2D.asset", ["filter", "iterations"]),
        "FlowMapStage.asset": asset("FlowMapStage.asset", ["iterations", "normMin", "normMax"]),
        "MeshTileStage.asset": asset("MeshTileStage.asset", ["meshType"]),
    }
    json.dump(params, open(os.path.join(HERE, "stage_params.json"), "w"), indent=1, sort_keys=True)
    print(f"wrote {len(out)} tables, {len(params)} assets")


if __name__ == "__main__":
    main()
