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
