"""Byte- and value-level diff of two PipelineState dumps (the reference's PipelineSerdeManager format, written by
unity/Editor/NoizeBench.cs on a Unity box, by tools/dump_chain.py / serde.py here).

    python tools/compare_dump.py <dir_a>/save__<alias_a> <dir_b>/save__<alias_b> [--tol 1e-5]
    python tools/compare_dump.py --log <Unity -logFile>        # collect the NOIZE_BENCH timing lines into a table

For every buffer present in both dumps (files.json ids): element count, byte equality, max |a-b|, the index of the worst
element, and how many elements differ at all / by more than `tol`.  Exit status 1 when a buffer differs by more than
`tol` (1e-5 of the normalised height range is the north-star tolerance; mesh indices and value erosion must be byte-equal).
This is the tool that turns "parity unpinned" into a number the day a Burst dump exists; nothing in this repository can
produce the Burst side.
"""
import argparse
import json
import os
import sys

import numpy as np

_DTYPES = {"Single": np.float32, "Int32": np.int32, "UInt32": np.uint32, "Double": np.float64, "Byte": np.uint8,
           "Int16": np.int16, "UInt16": np.uint16}


def load(save_dir):
    d = json.load(open(os.path.join(save_dir, "files.json")))
    out = {}
    for f in d.get("files", []):
        path = os.path.join(save_dir, "data", f["id"] + ".data")
        if os.path.exists(path):
            out[f["id"]] = np.fromfile(path, dtype=np.dtype(_DTYPES.get(f["type"], np.uint8)).newbyteorder("<"))
    return d, out


def compare(a_dir, b_dir, tol):
    da, a = load(a_dir)
    db, b = load(b_dir)
    print(f"A: {a_dir} (alias {da.get('alias')}, version {da.get('version')}, {len(a)} buffers)")
    print(f"B: {b_dir} (alias {db.get('alias')}, version {db.get('version')}, {len(b)} buffers)")
    worst = 0.0
    print(f"{'buffer':16s} {'elements':>12s} {'bytes equal':>12s} {'max |a-b|':>12s} {'at index':>10s} {'differ':>10s} {f'> {tol:g}':>10s}")
    for name in sorted(set(a) & set(b)):
        x, y = a[name], b[name]
        if x.size != y.size or x.dtype != y.dtype:
            print(f"{name:16s} size/type mismatch: {x.size} {x.dtype} vs {y.size} {y.dtype}")
            worst = float("inf")
            continue
        same = x.tobytes() == y.tobytes()
        if x.dtype.kind == "f":
            diff = np.abs(x.astype(np.float64) - y.astype(np.float64))
            diff[np.isnan(diff)] = np.inf
        else:
            diff = (x != y).astype(np.float64)
        k = int(diff.argmax()) if diff.size else 0
        m = float(diff[k]) if diff.size else 0.0
        worst = max(worst, m)
        print(f"{name:16s} {x.size:12d} {str(same):>12s} {m:12.3e} {k:10d} {int((diff > 0).sum()):10d} {int((diff > tol).sum()):10d}")
    for name in sorted(set(a) ^ set(b)):
        print(f"{name:16s} only in {'A' if name in a else 'B'}")
    return worst


def collect_log(path):
    rows = []
    for line in open(path, errors="replace"):
        i = line.find("NOIZE_BENCH {")
        if i >= 0:
            try:
                rows.append(json.loads(line[i + len("NOIZE_BENCH "):]))
            except json.JSONDecodeError:
                pass
    print(f"{'config':6s} {'stage':12s} {'arm':6s} {'ms best':>10s} {'ms mean':>10s} {'Mcells/s':>12s} {'workers':>8s}")
    for r in sorted(rows, key=lambda r: (r["config"], r["stage"], r["arm"])):
        print(f"{r['config']:6s} {r['stage']:12s} {r['arm']:6s} {r['ms_best']:10.3f} {r['ms_mean']:10.3f} {r['mcells_s_best']:12.1f} {r['job_workers']:8d}")
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dirs", nargs="*")
    ap.add_argument("--tol", type=float, default=1e-5)
    ap.add_argument("--log")
    args = ap.parse_args()
    if args.log:
        collect_log(args.log)
        return 0
    if len(args.dirs) != 2:
        ap.error("give two save__<alias> directories")
    return 1 if compare(args.dirs[0], args.dirs[1], args.tol) > args.tol else 0


if __name__ == "__main__":
    sys.exit(main())
