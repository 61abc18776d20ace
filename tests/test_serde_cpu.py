"""PipelineState on-disk format (Pipeline/PipelineState/PipelineSerialization.cs): layout, JSON shape, round trip."""
import json
import os

import numpy as np


def test_serde_layout_json_and_round_trip(tmp_path, nz):
    from noize_job_b200 import serde
    m = serde.PipelineSerdeManager(str(tmp_path), "world", "0.0.1")
    tile = np.random.default_rng(1).random((16, 16), dtype=np.float32)
    m.WriteData(tile, "TERRAIN_0_1")
    m.WriteData(np.arange(10, dtype=np.uint32), "idx/7")                         # '/' is an invalid file-name character
    base = tmp_path / "save__world"
    assert (base / "data" / "TERRAIN_0_1.data").read_bytes() == tile.tobytes()   # raw little-endian floats, no header
    assert (base / "data" / "idx_7.data").stat().st_size == 40
    d = json.loads((base / "files.json").read_text())
    assert list(d) == ["alias", "version", "files"] and d["alias"] == "world" and d["version"] == "0.0.1"
    assert d["files"] == [{"id": "TERRAIN_0_1", "type": "Single", "size": 256}, {"id": "idx/7", "type": "UInt32", "size": 10}]
    # a second manager reads the directory back (FileDirectory.FromFile) and the data (ReadData)
    m2 = serde.PipelineSerdeManager(str(tmp_path), "world", "ignored")
    assert m2.CachedSize("TERRAIN_0_1") == 256 and m2.CachedSize("missing") == -1 and m2.CachedSize("idx/7", np.uint32) == 10
    assert np.array_equal(m2.ReadData("TERRAIN_0_1").reshape(16, 16), tile)
    assert m2.ReadData("missing") is None
    m2.WriteData(tile[:8], "TERRAIN_0_1")                                        # SetCount updates in place
    assert json.loads((base / "files.json").read_text())["files"][0]["size"] == 128
    assert serde.clean_file_name("a//b/.c..") == "a_b_.c"


def test_compare_dump_reports_byte_and_value_differences(tmp_path, oracle):
    """tools/compare_dump.py: the tool that diffs a Unity-side (Burst) dump against ours, stage by stage."""
    import subprocess
    import sys
    from noize_job_b200 import serde
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tile = oracle.fractal(64, 64, 3, 0.4, octaves=13, noise_size=1700).reshape(-1)
    a = serde.PipelineSerdeManager(str(tmp_path), "burst_C1", "t")
    b = serde.PipelineSerdeManager(str(tmp_path), "gpu_C1", "t")
    a.WriteData(tile, "noise")
    b.WriteData(tile, "noise")
    a.WriteData(tile, "gauss5x17")
    off = tile.copy()
    off[100] += np.float32(3e-5)
    b.WriteData(off, "gauss5x17")
    cmd = [sys.executable, os.path.join(root, "tools", "compare_dump.py"), str(tmp_path / "save__burst_C1"), str(tmp_path / "save__gpu_C1")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1, r.stdout
    rows = {l.split()[0]: l.split() for l in r.stdout.splitlines() if l.startswith(("noise", "gauss5x17"))}
    assert rows["noise"][2] == "True" and float(rows["noise"][3]) == 0.0
    assert rows["gauss5x17"][2] == "False" and rows["gauss5x17"][4] == "100" and rows["gauss5x17"][5] == "1"
    assert subprocess.run(cmd + ["--tol", "1e-4"], stdout=subprocess.PIPE).returncode == 0
    log = tmp_path / "unity.log"
    log.write_text('x\nNOIZE_BENCH {"arm":"burst","config":"C1","stage":"noise","ms_best":1.5,"ms_mean":1.6,"mcells_s_best":43.7,"reps":5,"job_workers":15,"processors":16}\n')
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "compare_dump.py"), "--log", str(log)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "burst" in r.stdout and "43.7" in r.stdout
